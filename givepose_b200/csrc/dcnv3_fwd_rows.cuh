// givepose_b200 -- DCNv3 forward for 3x3 kernels with TMA-staged offset / mask rows and 16-byte sampling records (sm_100a).
//
// Same decomposition as dcnv3_fwd_tile (dcnv3_kernels.cuh: unit = (output pixel, group), CTA = tile_h x tile_w pixels x gs
// groups, L lanes x VEC channels per unit, one 128-byte line per corner gather) -- the reference is
// network/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:216-282 -- with the two non-gather costs of that kernel cut
// (ncu, profiles/r01_it6_ncu_summary.md: 55.6 L1 wavefronts per unit against 37 for the gathers + the store):
//
//   rows     the CTA's offset rows (tile_h x tile_w pixels x gs*18 values) and mask rows (gs*9) are two 3-D TMA boxes
//            {values, tile_w, tile_h} of the {G*P*2 | G*P, Wo, N*Ho} row tensors, issued by one elected thread and signalled on
//            an mbarrier: no LSU instructions and no per-thread 72-byte-stride global loads (7.6 wavefronts per unit before).
//            This is cuh:243-266's row reads, done once per CTA by the copy engine.
//   records  TWO threads per unit turn its 9 points into 16-byte records {byte offset | fx | fy << 1, a, b, lw}: the four
//            mask-folded bilinear weights of cuh:55-78 are rank one, w1..w4 = {a, b} x {1 - lw, lw} with a = hh * mask,
//            b = lh * mask.  Border handling stays branch-free: a clipped column sets fx = 0 (both column loads hit the one valid
//            column) and folds that column's weight into a and b; a clipped row sets fy = 0 and zeroes the missing row's weight;
//            an out-of-range sample (cuh:268-269) has a = b = 0 and borrows an address of its unit.  One LDS.128 per point and
//            unit instead of LDS.128 + LDS.64, a third less shared memory.
//   sampling unchanged: four unconditional 16-byte gathers + 16 FMAs per point and lane, no coordinate arithmetic.
#pragma once

#include <cuda.h>

#include "dcnv3_kernels.cuh"

namespace gp {

__device__ __forceinline__ uint32_t rows_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T> __device__ __forceinline__ void lds_pair(const T *p, float &a, float &b);
template <> __device__ __forceinline__ void lds_pair<float>(const float *p, float &a, float &b) {
    const float2 r = *reinterpret_cast<const float2 *>(p);
    a = r.x; b = r.y;
}
template <> __device__ __forceinline__ void lds_pair<__nv_bfloat16>(const __nv_bfloat16 *p, float &a, float &b) {
    const uint32_t r = *reinterpret_cast<const uint32_t *>(p);
    a = __uint_as_float(r << 16); b = __uint_as_float(r & 0xffff0000u);
}
template <> __device__ __forceinline__ void lds_pair<__half>(const __half *p, float &a, float &b) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(p));
    a = f.x; b = f.y;
}

// shared memory: [rows: offsets | masks] [records] [unit flags] [mbarrier]; bw_* = box widths in elements (16-byte multiples)
__host__ __device__ constexpr size_t fwd_rows_smem(int TP, int gs, int bw_off, int bw_msk, int es) {
    return 128 /*alignment slack*/ + (((size_t)TP * bw_off * es + 127) / 128) * 128 + (((size_t)TP * bw_msk * es + 127) / 128) * 128 +
           (size_t)TP * gs * 9 * 16 + (size_t)TP * gs * 4 + 16;
}

// 4 CTAs per SM (64 registers): budgets for 5 / 6 CTAs (48 / 40 registers, spills, less L1 next to the larger shared-memory
// carve-out) were measured at 0.534 / 0.612 ms against 0.475 ms (profiles/r02_it6_fwd_rows_occupancy.json)
template <typename T, int VEC, int L, bool SOFTMAX>
__global__ void __launch_bounds__(kTileThreads, kTileMinBlocks)
dcnv3_fwd_rows(const T *__restrict__ in, T *__restrict__ out, const __grid_constant__ CUtensorMap map_off,
               const __grid_constant__ CUtensorMap map_msk, const __grid_constant__ KParams p, int bw_off, int bw_msk) {
    extern __shared__ uint8_t smem_rows[];
    constexpr int P = 9;
    const TileCtx t = decode_tile(p);
    uint8_t *sbase = smem_rows + ((128u - (rows_smem_u32(smem_rows) & 127u)) & 127u);
    const size_t off_bytes = (((size_t)t.TP * bw_off * sizeof(T) + 127) / 128) * 128;
    const size_t msk_bytes = (((size_t)t.TP * bw_msk * sizeof(T) + 127) / 128) * 128;
    const T *s_off = reinterpret_cast<const T *>(sbase);
    const T *s_msk = reinterpret_cast<const T *>(sbase + off_bytes);
    float4 *s_rec = reinterpret_cast<float4 *>(sbase + off_bytes + msk_bytes);
    unsigned *s_unit = reinterpret_cast<unsigned *>(s_rec + t.n_ul * P);
    const uint32_t bar = (rows_smem_u32(s_unit + t.n_ul) + 7u) & ~7u;

    // TMA moves 16-byte units: a box starts at the 16-byte boundary at or below the CTA's first value, `lead` values early
    // (the box widths leave room for it)
    constexpr int PER16 = 16 / (int)sizeof(T);
    const int lead_off = (t.g0 * (P * 2)) & (PER16 - 1), lead_msk = (t.g0 * P) & (PER16 - 1);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)((size_t)t.TP * (bw_off + bw_msk) * sizeof(T));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const int row0 = t.b * p.Ho + t.oh0;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(rows_smem_u32(s_off)), "l"(&map_off), "r"(bar), "r"(t.g0 * (P * 2) - lead_off), "r"(t.ow0), "r"(row0) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(rows_smem_u32(s_msk)), "l"(&map_msk), "r"(bar), "r"(t.g0 * P - lead_msk), "r"(t.ow0), "r"(row0) : "memory");
    }
    __syncthreads();   // the barrier is initialised before anyone polls it
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "ROWS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra ROWS_DONE;\n\t"
        "bra ROWS_WAIT;\n\t"
        "ROWS_DONE:\n\t"
        "}" ::"r"(bar) : "memory");

    // ---- records: two threads per unit (points 0..4 / 5..8) -----------------------------------------------------------------
    const int C = p.C, WC = p.W * C;
    {
        const int sub = threadIdx.x & 1;
        const int upi = blockDim.x >> 1;
        for (int e0 = 0; e0 < t.n_ul; e0 += upi) {
            const int e = e0 + (threadIdx.x >> 1);   // (pix, g_local) with g_local fastest: the order of the staged rows
            const int pix = e >> p.lg_gs, gl = e & (p.gs - 1);
            const int oh = t.oh0 + (pix >> p.lg_tw), ow = t.ow0 + (pix & (p.tile_w - 1));
            const int ul = gl * t.TP + pix;
            const bool live = e < t.n_ul && oh < p.Ho && ow < p.Wo;   // tile overhang: the sampling loop never visits the unit
            const T *orow = s_off + lead_off + (live ? pix * bw_off + gl * (P * 2) : 0), *mrow = s_msk + lead_msk + (live ? pix * bw_msk + gl * P : 0);
            const float p0_h_ = origin<float>(p.base_h + oh * p.sh, p.half_h, p.scale);
            const float p0_w_ = origin<float>(p.base_w + ow * p.sw, p.half_w, p.scale);
            float mx = 0.f, inv = 1.f;
            if (SOFTMAX) {   // softmax over the 9 logits of the (pixel, group) row: modules/dcnv3.py:332-333
                mx = to_acc<T>(mrow[0]);
#pragma unroll
                for (int i = 1; i < P; ++i) mx = fmaxf(mx, to_acc<T>(mrow[i]));
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < P; ++i) sum += expf(to_acc<T>(mrow[i]) - mx);
                inv = 1.f / sum;
            }
            float4 rec[5];
            unsigned out_mask = 0u;
            int ubase = -1;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int pt = sub * 5 + j;   // p = i*kh + j', kernel WIDTH index i slow (cuh:257-258); pt 9 is a pad slot
                const int ptc = pt < P ? pt : P - 1;
                float ox, oy;
                lds_pair<T>(orow + 2 * ptc, ox, oy);   // (w, h) pair, cuh:261-262
                float m = to_acc<T>(mrow[ptc]);
                if (SOFTMAX) m = expf(m - mx) * inv;
                Point<float> sp;
                locate<float>(sp, p0_h_, p0_w_, (ptc % 3) * p.dh, (ptc / 3) * p.dw, ox, oy, p.scale, p.H, p.W);
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sp.flags & F_IN) {
                    const bool hl = sp.h_low >= 0, wl = sp.w_low >= 0;
                    const bool hh_ok = sp.h_low + 1 <= p.H - 1, wh_ok = sp.w_low + 1 <= p.W - 1;
                    const int r0 = hl ? sp.h_low : sp.h_low + 1, c0 = wl ? sp.w_low : sp.w_low + 1;
                    const unsigned ob = (unsigned)((r0 * WC + c0 * C) * (int)sizeof(T));   // a multiple of 16 (VEC * sizeof(T) == 16 ...)
                    float a = hl ? sp.hh * m : 0.f, b = hh_ok ? sp.lh * m : 0.f, lw = sp.lw;   // rows outside contribute 0 (cuh:55-75)
                    if (!wl) { a *= sp.lw; b *= sp.lw; lw = 0.f; }            // only column w_low + 1 is inside
                    else if (!wh_ok) { a *= sp.hw; b *= sp.hw; lw = 0.f; }    // only column w_low is inside
                    r = make_float4(__uint_as_float(ob | (wl && wh_ok ? 1u : 0u) | (hl && hh_ok ? 2u : 0u)), a, b, lw);
                    if (ubase < 0) ubase = (int)ob;
                } else {
                    out_mask |= 1u << j;
                }
                rec[j] = r;
            }
            // out-of-range samples borrow an address their unit reads anyway (weights 0); a unit without any in-range sample is dead
            const int other = __shfl_xor_sync(0xffffffffu, ubase, 1);
            if (ubase < 0) ubase = other;
            if (live) {
                float4 *dst = s_rec + ul * P + sub * 5;
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (sub * 5 + j < P) {
                        float4 r = rec[j];
                        if ((out_mask >> j) & 1u) r.x = __uint_as_float(ubase >= 0 ? (unsigned)ubase : 0u);
                        dst[j] = r;
                    }
                if (sub == 0) s_unit[ul] = ubase >= 0 ? 1u : 0u;
            } else if (e < t.n_ul && sub == 0) {
                s_unit[ul] = 0u;
            }
        }
    }
    __syncthreads();

    // ---- sampling ---------------------------------------------------------------------------------------------------------------
    const int cl = threadIdx.x % L;
    const unsigned Cb = (unsigned)(C * (int)sizeof(T)), WCb = (unsigned)(WC * (int)sizeof(T));
    const T *in_b = in + (long long)t.b * p.H * WC + cl * VEC;
    constexpr int UPB = kTileThreads / L;
    const int n_pass = (t.n_ul + UPB - 1) / UPB;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int ul = pass * UPB + threadIdx.x / L;
        const UnitPos u = unit_pos(ul, p, t);
        if (!(ul < t.n_ul && u.oh < p.Ho && u.ow < p.Wo)) continue;
        const int g = t.g0 + u.gl;
        const char *in_g = reinterpret_cast<const char *>(in_b + g * p.gc);
        const float4 *rw = s_rec + ul * P;
        float acc[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
        if (s_unit[ul]) {   // a unit whose samples are all out of range reads nothing and writes zeros (cuh:268-269)
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const float4 r = rw[k];
                const unsigned o = __float_as_uint(r.x);
                const unsigned fx = o & 1u, fy = (o >> 1) & 1u;
                const char *p1 = in_g + (o & ~3u);
                const char *p2 = p1 + (size_t)(fx * Cb);
                const char *p3 = p1 + (size_t)(fy * WCb);
                const char *p4 = p3 + (size_t)(fx * Cb);
                float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p1), v1);
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p2), v2);
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p3), v3);
                Vec<T, VEC>::load(reinterpret_cast<const T *>(p4), v4);
                const float hw = 1.f - r.w;
                const float w1 = r.y * hw, w2 = r.y * r.w, w3 = r.z * hw, w4 = r.z * r.w;   // (hh, lh) * mask x (hw, lw): cuh:76-78
#pragma unroll
                for (int c = 0; c < VEC; ++c)
                    acc[c] = fmaf(w1, v1[c], fmaf(w2, v2[c], fmaf(w3, v3[c], fmaf(w4, v4[c], acc[c]))));
            }
        }
        const long long q = ((long long)t.b * p.Ho + u.oh) * p.Wo + u.ow;
        Vec<T, VEC>::store_stream(out + q * C + g * p.gc + cl * VEC, acc);
    }
}

}  // namespace gp
