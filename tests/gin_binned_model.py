"""NumPy model of ``dcnv3_gin_binned`` (givepose_b200/csrc/dcnv3_gin_binned.cuh), step for step: window of footprint cells,
counting sort by cell, 2x2 destination blocks fed by the contiguous run of each of the three cell rows, flush with
the per-pixel bounds check.  Test infrastructure: it lets the CPU suite check the kernel's *algorithm* (block / run /
weight bookkeeping) against the C oracle without a GPU; the kernel itself is checked on the GPU by tests/test_dcnv3_gpu.py.
"""
import numpy as np

CELLS, WIN_W = 1024, 32   # kGinCells, kGinWinW


def grad_input_binned(off, msk, gout, N, H, W, G, gc, kh, kw, sh, sw, ph, pw, dh, dw, scale, rc, Ho, Wo, hw_low, flags,
                      tile=(8, 8), stats=None, row_walk=False):
    """off (rows, G*P*2), msk (rows, G*P), gout (N,Ho,Wo,G*gc) float arrays; hw_low (n,2) / flags (n,) from the oracle's
    index() (the integer part of the contract).  Returns grad_input (N,H,W,G*gc) float64.
    row_walk=True models dcnv3_bwd_fused.cuh instead (768-cell window, one visit per sample along each cell row, two
    column slots -- even / odd destination column -- flushed when the row's column moves on)."""
    cells_max = 768 if row_walk else CELLS
    P = kh * kw - rc
    C = G * gc
    gin = np.zeros((N, H, W, C))
    off = off.reshape(-1)[: N * Ho * Wo * G * P * 2].reshape(N, Ho, Wo, G, P, 2).astype(np.float64)
    msk = msk.reshape(-1)[: N * Ho * Wo * G * P].reshape(N, Ho, Wo, G, P).astype(np.float64)
    hw_low = hw_low.reshape(N, Ho, Wo, G, P, 2)
    flags = flags.reshape(N, Ho, Wo, G, P)
    cidx = (kw // 2) * kh + kh // 2
    th, tw = tile
    half_h, half_w = (dh * (kh - 1)) >> 1, (dw * (kw - 1)) >> 1
    for b in range(N):
        for oh0 in range(0, Ho, th):
            for ow0 in range(0, Wo, tw):
                for g in range(G):
                    samples = []   # (h_low, w_low, lh, lw, m, grad_output row)
                    for oh in range(oh0, min(oh0 + th, Ho)):
                        for ow in range(ow0, min(ow0 + tw, Wo)):
                            p0h = (half_h - ph + oh * sh) - half_h * scale
                            p0w = (half_w - pw + ow * sw) - half_w * scale
                            for pt in range(P):
                                if not flags[b, oh, ow, g, pt] & 1:
                                    continue
                                kk = pt + (1 if rc and pt >= cidx else 0)
                                i, j = kk // kh, kk % kh
                                loc_w = p0w + (i * dw + off[b, oh, ow, g, pt, 0]) * scale
                                loc_h = p0h + (j * dh + off[b, oh, ow, g, pt, 1]) * scale
                                hl, wl = int(hw_low[b, oh, ow, g, pt, 0]), int(hw_low[b, oh, ow, g, pt, 1])
                                samples.append((hl, wl, loc_h - hl, loc_w - wl, msk[b, oh, ow, g, pt],
                                                gout[b, oh, ow, g * gc:(g + 1) * gc].astype(np.float64)))
                    if not samples:
                        continue
                    mn_h, mx_h = min(s[0] for s in samples), max(s[0] for s in samples)
                    mn_w, mx_w = min(s[1] for s in samples), max(s[1] for s in samples)
                    WWa = min(mx_w - mn_w + 1, WIN_W)
                    WHa = min(mx_h - mn_h + 1, cells_max // WWa)
                    NCa = WHa * WWa
                    bins = [[] for _ in range(NCa)]
                    for s in samples:
                        r, c = s[0] - mn_h, s[1] - mn_w
                        if r < WHa and c < WWa:
                            bins[r * WWa + c].append((s[2], s[3], s[4], c, s[5]))
                        else:   # outside the window: plain per-corner scatter
                            hl, wl, lh, lw, m, gv = s
                            for (yy, xx, wgt) in ((hl, wl, (1 - lh) * (1 - lw)), (hl, wl + 1, (1 - lh) * lw),
                                                  (hl + 1, wl, lh * (1 - lw)), (hl + 1, wl + 1, lh * lw)):
                                if 0 <= yy < H and 0 <= xx < W:
                                    gin[b, yy, xx, g * gc:(g + 1) * gc] += wgt * m * gv
                            if stats is not None:
                                stats["overflow"] = stats.get("overflow", 0) + 1
                    start = np.concatenate([[0], np.cumsum([len(x) for x in bins])])
                    srt = [rec for cell in bins for rec in cell]
                    if row_walk:
                        for row in range(WHa):
                            i, end = start[row * WWa], start[(row + 1) * WWa]
                            if i == end:
                                continue
                            y = mn_h + row
                            acc = {0: [np.zeros(gc), np.zeros(gc)], 1: [np.zeros(gc), np.zeros(gc)]}   # slot -> [top, bottom]
                            cur = {0: -2, 1: -1}

                            def flush(slot):
                                c = cur[slot]
                                x = mn_w + c
                                if c < 0 or x < 0 or x >= W:
                                    return
                                for dy in (0, 1):
                                    if 0 <= y + dy < H and np.any(acc[slot][dy] != 0):
                                        gin[b, y + dy, x, g * gc:(g + 1) * gc] += acc[slot][dy]
                                        if stats is not None:
                                            stats["flush_lines"] = stats.get("flush_lines", 0) + 1
                            for k in range(i, end):
                                lh, lw, m, col, gv = srt[k]
                                odd = col & 1
                                d = {0: col + odd, 1: col + 1 - odd}
                                for slot in (0, 1):
                                    if d[slot] != cur[slot]:
                                        flush(slot)
                                        acc[slot] = [np.zeros(gc), np.zeros(gc)]
                                        cur[slot] = d[slot]
                                hw = 1 - lw
                                am, bm = (1 - lh) * m, lh * m
                                w = {0: lw if odd else hw, 1: hw if odd else lw}
                                for slot in (0, 1):
                                    acc[slot][0] += am * w[slot] * gv
                                    acc[slot][1] += bm * w[slot] * gv
                            flush(0)
                            flush(1)
                        if stats is not None:
                            stats["corner_lines"] = stats.get("corner_lines", 0) + 4 * len(samples)
                        continue
                    BW, BH = (WWa + 2) >> 1, (WHa + 2) >> 1
                    for blk in range(BH * BW):
                        bi, bj = divmod(blk, BW)
                        d0, e0 = 2 * bi, 2 * bj
                        c_lo, c_hi = max(e0 - 1, 0), min(e0 + 1, WWa - 1)
                        acc = np.zeros((2, 2, gc))
                        visited = 0
                        for rr in (-1, 0, 1):
                            r = d0 + rr
                            if r < 0 or r >= WHa:
                                continue
                            beg, end = start[r * WWa + c_lo], start[r * WWa + c_hi + 1]
                            visited += end - beg
                            for i in range(beg, end):
                                lh, lw, m, c, gv = srt[i]
                                dx = c - e0
                                hw = 1 - lw
                                cw0 = hw if dx == 0 else (lw if dx < 0 else 0.0)
                                cw1 = lw if dx == 0 else (hw if dx > 0 else 0.0)
                                if rr <= 0:
                                    a = ((1 - lh) if rr == 0 else lh) * m
                                    acc[0, 0] += a * cw0 * gv
                                    acc[0, 1] += a * cw1 * gv
                                if rr >= 0:
                                    a = (lh if rr == 0 else (1 - lh)) * m
                                    acc[1, 0] += a * cw0 * gv
                                    acc[1, 1] += a * cw1 * gv
                        if not visited:
                            continue
                        for dy in (0, 1):
                            for dxx in (0, 1):
                                y, x = mn_h + d0 + dy, mn_w + e0 + dxx
                                if 0 <= y < H and 0 <= x < W and np.any(acc[dy, dxx] != 0):
                                    gin[b, y, x, g * gc:(g + 1) * gc] += acc[dy, dxx]
                                    if stats is not None:
                                        stats["flush_lines"] = stats.get("flush_lines", 0) + 1
                    if stats is not None:
                        stats["corner_lines"] = stats.get("corner_lines", 0) + 4 * len(samples)
    return gin
