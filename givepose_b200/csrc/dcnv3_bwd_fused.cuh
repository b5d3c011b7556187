// givepose_b200 -- DCNv3 backward, ONE kernel with in-SM pre-aggregation of grad_input (sm_100a; GP_OPT_BWD_MODE 2).
//
// Measured background (profiles/r02_*): the one-pass scatter backward (dcnv3_bwd_tile, mode 0) issues 36 line reductions
// per (pixel, group) unit = 8.2 GB of fp32 reductions per launch at config 2 and runs AT the measured ceiling of that
// pattern (tools/micro/gather_rates.cu mode 5/6: 1.44 ms for the reductions alone, the kernel takes 1.38 ms).  Splitting
// the backward into a gather kernel + a binned grad_input kernel (mode 1, dcnv3_gin_binned.cuh) cuts the reduction
// sectors 7.6x but costs 0.86 + 1.0 ms: the two halves cannot overlap.  This kernel keeps them in one CTA, so the
// reduction-free gather phase (L1 / latency bound) of one CTA overlaps the sort + accumulate phase (issue bound) of its
// neighbours on the SM, and each sample is visited ONCE:
//
//   records  = build_records<BWD> as in dcnv3_bwd_tile (one thread per unit), the footprint cell (h_low, w_low) packed
//              into the record's spare word; threads that build nothing stage the tile's grad_output rows as fp32.
//   sort     = window of cells from two warp reductions, counting sort of the CTA's samples by cell (native integer
//              ATOMS.ADD ranks, one block scan): samples of one cell ROW end up contiguous, ordered by column.
//   row walk = a group of L = gc/4 lanes walks one cell row left to right with FOUR register accumulators: destination
//              rows (h_low, h_low + 1) x the even / odd destination column currently open.  A sample adds its four
//              corner contributions with 16 FMAs; when the row's column moves on, the closed column is flushed with ONE
//              red.global.add.v4.f32 per lane per destination pixel (corners outside the image are dropped there: the
//              per-corner bounds check of cuh:116-140).  Every destination pixel of the window is reduced at most twice
//              (from cell row y and cell row y - 1): ~12 line reductions per unit instead of 36 at the reference's test
//              distribution, fewer for model-like offsets.  Samples outside the 32-column / 768-cell window take a
//              scalar atomicAdd path.
//   gathers  = the grad_offset / grad_mask half, identical to dcnv3_bwd_tile<GIN = false>: 36 corner gathers, four dot
//              products per point, transposing-butterfly reduction over the unit's lanes, results parked in the consumed
//              records and written back coalesced.
#pragma once

#include "dcnv3_gin_binned.cuh"

namespace gp {

constexpr int kFusedCells = 768;   // counting-sort bins per CTA
constexpr int kFusedThreads = 128;
constexpr int kFusedSPT = 5;       // samples per thread in the sort (tile * P <= 640)

__host__ __device__ constexpr size_t bwd_fused_smem(int TP, int P, int L) {
    //      records (16 + 8 bytes)      s_unit          fp32 grad_output rows   bins             sorted entries        per-warp scratch
    return (size_t)TP * P * 24 + (size_t)TP * 4 + (size_t)TP * L * 16 + (kFusedCells + 1) * 4 + (size_t)TP * P * 4 + 64 * 4 + 16;
}

template <typename T, int L, bool P9, int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB)
dcnv3_bwd_fused(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk,
                const T *__restrict__ gout, float *__restrict__ gin, T *__restrict__ goff, T *__restrict__ gmsk,
                const __grid_constant__ KParams p) {
    constexpr int NT = kFusedThreads, NW = NT / 32, VEC = 4, SPT = kFusedSPT;
    extern __shared__ float4 smem4[];
    const TileCtx t = decode_tile(p);   // gs == 1
    const int P = P9 ? 9 : p.P;
    const int n_rec = t.n_ul * P;
    float4 *s_w = smem4;                                              // [n_rec] {lh, lw, mask, cell}
    float4 *s_g4 = s_w + n_rec;                                       // [TP][L] fp32 grad_output rows
    int2 *s_bf = reinterpret_cast<int2 *>(s_g4 + t.TP * L);           // [n_rec] {corner-1 byte offset, flags}
    unsigned *s_unit = reinterpret_cast<unsigned *>(s_bf + n_rec);    // [TP]
    int *s_A = reinterpret_cast<int *>(s_unit + t.TP);                // [kFusedCells + 1]
    unsigned *s_sorted = reinterpret_cast<unsigned *>(s_A + kFusedCells + 1);   // [n_rec] record | unit << 11 | column << 19
    int *s_scr = reinterpret_cast<int *>(s_sorted + n_rec);           // [64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = t.g0, C = p.C, WC = p.W * C;

    // ---- grad_output rows of the tile -> shared memory (fp32); zero the bins ----------------------------------------
    for (int i = tid; i <= kFusedCells; i += NT) s_A[i] = 0;
    for (int i = tid; i < t.TP * L; i += NT) {
        const int ul = i / L, c4 = i - ul * L;
        const int oh = t.oh0 + (ul >> p.lg_tw), ow = t.ow0 + (ul & (p.tile_w - 1));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oh < p.Ho && ow < p.Wo) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            v = load4_f32<T>(gout + q * C + g * p.gc + c4 * 4);
        }
        s_g4[i] = v;
    }
    build_records<T, false, true, P9>(off, msk, s_w, s_bf, s_unit, p, t);   // ends with __syncthreads()

    const long long img = (long long)t.b * p.H * WC;
    float *gin_g = gin + img + g * p.gc;

    // ---- window of footprint cells ---------------------------------------------------------------------------------------
    int s_key[SPT];   // (h_low + 1) << 16 | (w_low + 1), or -1
    int mn_h = 0x7fffffff, mx_h = -0x7fffffff, mn_w = 0x7fffffff, mx_w = -0x7fffffff;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int r = tid + k * NT;
        s_key[k] = -1;
        bool live = r < n_rec;
        if (live) {   // build_records leaves the records of tile-overhang units unwritten: never look at them
            const int ul = P9 ? r / 9 : r / P;
            live = t.oh0 + (ul >> p.lg_tw) < p.Ho && t.ow0 + (ul & (p.tile_w - 1)) < p.Wo;
        }
        if (live && s_bf[r].y != 0) {   // flags != 0 <=> the sample is in range (cuh:268-269)
            const int key = __float_as_int(s_w[r].w);
            s_key[k] = key;
            const int h = (key >> 16) - 1, w = (key & 0xffff) - 1;
            mn_h = min(mn_h, h); mx_h = max(mx_h, h); mn_w = min(mn_w, w); mx_w = max(mx_w, w);
        }
    }
    mn_h = warp_min(mn_h); mx_h = warp_max(mx_h); mn_w = warp_min(mn_w); mx_w = warp_max(mx_w);
    if (lane == 0) { s_scr[warp * 4 + 0] = mn_h; s_scr[warp * 4 + 1] = mx_h; s_scr[warp * 4 + 2] = mn_w; s_scr[warp * 4 + 3] = mx_w; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        mn_h = min(mn_h, s_scr[w * 4 + 0]); mx_h = max(mx_h, s_scr[w * 4 + 1]);
        mn_w = min(mn_w, s_scr[w * 4 + 2]); mx_w = max(mx_w, s_scr[w * 4 + 3]);
    }
    const bool any = mx_h >= mn_h;   // uniform over the CTA
    const int WWa = any ? min(mx_w - mn_w + 1, kGinWinW) : 1;
    const int WHa = any ? min(mx_h - mn_h + 1, kFusedCells / WWa) : 0;
    const int NCa = WHa * WWa;

    // ---- histogram; samples outside the window scatter directly ------------------------------------------------------------
    int s_cell[SPT], s_rank[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        s_cell[k] = -1;
        s_rank[k] = 0;
        if (s_key[k] >= 0) {
            const int rr = tid + k * NT;
            const int h_low = (s_key[k] >> 16) - 1, w_low = (s_key[k] & 0xffff) - 1;
            const int r = h_low - mn_h, c = w_low - mn_w;
            if (r < WHa && c < WWa) {
                s_cell[k] = r * WWa + c;
                s_rank[k] = atomicAdd(&s_A[s_cell[k]], 1);
            } else {
                const float4 rec = s_w[rr];
                const float lh = rec.x, lw = rec.y, hh = 1.f - lh, hw = 1.f - lw, m = rec.z;
                const unsigned fl = (unsigned)s_bf[rr].y;
                const int ul = P9 ? rr / 9 : rr / P;
                float *b1 = gin_g + ((long long)h_low * p.W + w_low) * C;
                const float *gr = reinterpret_cast<const float *>(s_g4 + ul * L);
                for (int ch = 0; ch < p.gc; ++ch) {
                    const float tg = gr[ch] * m;
                    if (fl & F_C1) atomicAdd(b1 + ch, hh * hw * tg);
                    if (fl & F_C2) atomicAdd(b1 + C + ch, hh * lw * tg);
                    if (fl & F_C3) atomicAdd(b1 + WC + ch, lh * hw * tg);
                    if (fl & F_C4) atomicAdd(b1 + WC + C + ch, lh * lw * tg);
                }
            }
        }
    }
    __syncthreads();

    // ---- exclusive scan of the bins (s_A[NCa] = number of binned samples) -------------------------------------------------
    {
        const int ipt = (NCa + NT - 1) / NT;
        const int base = tid * ipt;
        int sum = 0;
        for (int k = 0; k < ipt; ++k)
            if (base + k < NCa) sum += s_A[base + k];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_scr[32 + warp] = incl;
        __syncthreads();
        int excl = incl - sum;
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (w < warp) excl += s_scr[32 + w];
        for (int k = 0; k < ipt; ++k)
            if (base + k < NCa) {
                const int c = s_A[base + k];
                s_A[base + k] = excl;
                excl += c;
            }
        if (ipt > 0 && base < NCa && NCa <= base + ipt) s_A[NCa] = excl;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SPT; ++k)
        if (s_cell[k] >= 0) {
            const int rr = tid + k * NT;
            const int ul = P9 ? rr / 9 : rr / P;
            const int c = s_cell[k] % WWa;
            s_sorted[s_A[s_cell[k]] + s_rank[k]] = (unsigned)rr | ((unsigned)ul << 11) | ((unsigned)c << 19);
        }
    __syncthreads();

    // ---- row walk: grad_input of the window, one visit per sample, flushed per closed destination column -----------------
    {
        constexpr int NG = NT / L;
        const int gid = tid / L, cl = tid - gid * L;
        for (int row = gid; row < WHa; row += NG) {
            int i = s_A[row * WWa];
            const int end = s_A[(row + 1) * WWa];
            if (i == end) continue;
            const int y = mn_h + row;   // cell row: destination rows y (corner weight hh) and y + 1 (lh)
            const bool top_ok = y >= 0 && y < p.H, bot_ok = y + 1 >= 0 && y + 1 < p.H;
            float *row_t = gin_g + (long long)y * WC + cl * 4, *row_b = row_t + WC;
            float4 a0t = make_float4(0.f, 0.f, 0.f, 0.f), a0b = a0t, a1t = a0t, a1b = a0t;
            int c0 = -2, c1 = -1;       // destination column (window relative) open in slot 0 (even columns) / slot 1 (odd)
            auto flush = [&](const float4 at, const float4 ab, int c) {
                const int x = mn_w + c;
                if (c < 0 || x < 0 || x >= p.W) return;
                if (top_ok && (at.x != 0.f || at.y != 0.f || at.z != 0.f || at.w != 0.f)) red_add_v4(row_t + (long long)x * C, at.x, at.y, at.z, at.w);
                if (bot_ok && (ab.x != 0.f || ab.y != 0.f || ab.z != 0.f || ab.w != 0.f)) red_add_v4(row_b + (long long)x * C, ab.x, ab.y, ab.z, ab.w);
            };
            for (; i < end; ++i) {
                const unsigned e = s_sorted[i];
                const int col = (int)(e >> 19);
                const float4 rec = s_w[e & 0x7ffu];
                const float4 gv = s_g4[((e >> 11) & 0xffu) * L + cl];
                const int odd = col & 1;
                const int d0 = col + odd, d1 = col + 1 - odd;   // the even / odd one of the sample's two destination columns
                if (d0 != c0) { flush(a0t, a0b, c0); a0t = a0b = make_float4(0.f, 0.f, 0.f, 0.f); c0 = d0; }
                if (d1 != c1) { flush(a1t, a1b, c1); a1t = a1b = make_float4(0.f, 0.f, 0.f, 0.f); c1 = d1; }
                const float lh = rec.x, lw = rec.y, m = rec.z, hw = 1.f - lw;
                const float am = (1.f - lh) * m, bm = lh * m;   // corner rows h_low (hh) / h_low + 1 (lh), times mask: cuh:116-140
                const float w0 = odd ? lw : hw, w1 = odd ? hw : lw;   // column w_low -> hw, w_low + 1 -> lw
                fma4(a0t, am * w0, gv);
                fma4(a0b, bm * w0, gv);
                fma4(a1t, am * w1, gv);
                fma4(a1b, bm * w1, gv);
            }
            flush(a0t, a0b, c0);
            flush(a1t, a1b, c1);
        }
    }
    // the row walk only READS the records; the gather phase below overwrites them unit by unit
    __syncthreads();

    // ---- gathers: grad_offset / grad_mask (dcnv3_bwd_tile<GIN = false>, top_grad from shared memory) --------------------
    const int cl = tid % L;
    const int Cb = C * (int)sizeof(T), WCb = WC * (int)sizeof(T);
    const T *in_b = in + img + cl * VEC;
    constexpr int UPB = NT / L;
    const int n_pass = (t.n_ul + UPB - 1) / UPB;
    const unsigned full = 0xffffffffu;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int ul = pass * UPB + tid / L;
        const UnitPos u = unit_pos(ul, p, t);
        const bool valid = ul < t.n_ul && u.oh < p.Ho && u.ow < p.Wo;
        if (!__any_sync(full, valid)) continue;   // warp-uniform
        const int ulc = valid ? ul : 0;
        const char *in_g = reinterpret_cast<const char *>(in_b + g * p.gc);
        float4 *rw = s_w + ulc * P;
        const int2 *rb = s_bf + ulc * P;
        float go[VEC] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {
            const float4 gv = s_g4[ulc * L + cl];
            go[0] = gv.x; go[1] = gv.y; go[2] = gv.z; go[3] = gv.w;
        }
        auto point_math = [&](const float4 r, const float (&v1)[VEC], const float (&v2)[VEC], const float (&v3)[VEC],
                              const float (&v4)[VEC], float &s_m, float &s_w_, float &s_h) {
            const float lh = r.x, lw = r.y, m = r.z;
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            float d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                d1 = fmaf(go[c], v1[c], d1);
                d2 = fmaf(go[c], v2[c], d2);
                d3 = fmaf(go[c], v3[c], d3);
                d4 = fmaf(go[c], v4[c], d4);
            }
            s_m = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;      // cuh:144
            s_w_ = m * (hh * (d2 - d1) + lh * (d4 - d3));     // cuh:145
            s_h = m * (hw * (d3 - d1) + lw * (d4 - d2));      // cuh:146
        };
        auto park = [&](int k, float s_m, float s_w_, float s_h) {
            float *slot = reinterpret_cast<float *>(rw + k);
            if (L == 1) {
                slot[0] = p.scale * s_w_; slot[1] = p.scale * s_h; slot[2] = s_m;
            } else if (L == 2) {
                if (cl == 0) { slot[2] = s_m; slot[0] = p.scale * s_w_; } else { slot[1] = p.scale * s_h; }
            } else {
                if (cl == 0) slot[2] = s_m;
                else if (cl == 2) slot[0] = p.scale * s_w_;
                else if (cl == 1) slot[1] = p.scale * s_h;
            }
        };
        if (P9 && __all_sync(full, valid && s_unit[ulc])) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int base = rb[k].x;
                const float4 r = rw[k];
                const char *p1 = in_g + base, *p3 = p1 + WCb;
                float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p1), v1);
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p1 + Cb), v2);
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p3), v3);
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p3 + Cb), v4);
                float s_m, s_w_, s_h;
                point_math(r, v1, v2, v3, v4, s_m, s_w_, s_h);
                unit_reduce3<L>(s_m, s_w_, s_h, cl);
                __syncwarp();
                park(k, s_m, s_w_, s_h);
            }
        } else {
            for (int k = 0; k < P; ++k) {
                const int2 bf = rb[k];
                const unsigned flags = valid ? (unsigned)bf.y : 0u;
                float s_m = 0.f, s_w_ = 0.f, s_h = 0.f;
                if (flags) {
                    const float4 r = rw[k];
                    float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
                    gather4<T, VEC>(in_g + bf.x, Cb, WCb, flags, v1, v2, v3, v4);
                    point_math(r, v1, v2, v3, v4, s_m, s_w_, s_h);
                }
                unit_reduce3<L>(s_m, s_w_, s_h, cl);
                __syncwarp();
                if (valid) park(k, s_m, s_w_, s_h);
            }
        }
    }
    __syncthreads();

    // ---- coalesced write-back of grad_offset / grad_mask -------------------------------------------------------------------
    for (int e = tid; e < n_rec; e += NT) {
        const int pix = e / P, pt = e - pix * P;
        const int oh = t.oh0 + (pix >> p.lg_tw), ow = t.ow0 + (pix & (p.tile_w - 1));
        if (oh < p.Ho && ow < p.Wo) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            const long long k = ((q * p.G + g) * (long long)P) + pt;
            const float4 res = s_w[e];   // out-of-range samples parked 0, 0, 0 (cuh:347-355)
            store_pair<T>(goff + 2 * k, res.x, res.y);
            gmsk[k] = from_acc<T, float>(res.z);
        }
    }
}

}  // namespace gp
