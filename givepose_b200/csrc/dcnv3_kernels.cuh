// givepose_b200 -- DCNv3 deformable-sampling kernels for sm_100a (forward, backward, index hook).
//
// Replaces the reference device code network/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh
//   forward  :216-282 (one thread per output SCALAR, offsets/mask re-read per channel, scalar gathers)
//   backward :285-888 (block = group_channels threads, 9x block-wide smem reductions, scalar atomics)
// with a different decomposition (nothing is translated):
//
//   unit      = one (output pixel q, group g): P sampling points, gc channels.
//   CTA       = a tile_h x tile_w patch of output pixels x `gs` groups of ONE image, so the input rows it
//               gathers form a compact window that stays in L1 (channel-last rows of one group are
//               contiguous: gc*sizeof(T) bytes = one 128-byte line for gc=32 fp32).
//   records   = ONE thread per unit reads the unit's offset / mask row (P (x, y) pairs + P mask values, straight from global
//               memory with read-only loads: a 72-byte stride between the threads of a warp) and turns its P points into
//               24-byte sampling records in shared memory: byte offset of the footprint + bounds flags, and either the
//               four mask-folded bilinear weights (forward) or lh / lw / mask (backward); unit-level arithmetic (pixel
//               decode, p0, the softmax of the fused variant) is done once per unit, kernel indices are compile-time for
//               3x3.  (An earlier one-thread-per-(unit, point) builder with coalesced row reads spent 35 % of the kernel's
//               instructions; staging the rows through shared memory was measured and dropped)  The
//               sampling loop reads a record with two broadcast LDS and does no coordinate arithmetic at all: the kernels
//               are bound by L1 wavefronts (one 128-byte line per corner per unit), so the instruction stream has to stay
//               well under 4 issue slots per wavefront.
//   thread    = VEC channels of one unit; L = gc/VEC consecutive lanes form a unit, so every corner
//               gather of a unit is one fully-used coalesced request.
//   backward  = same tiling; grad_input goes out as 16-byte vector reductions (REDG.ADD.F32x4) that
//               resolve in L2; grad_offset / grad_mask only need the four dot products
//               d_k = sum_c top_grad[c]*corner_k[c], which are combined per lane and summed over the L lanes
//               of the unit with a transposing butterfly (log2(L)+1 shuffles, no block barriers), parked in
//               the consumed record and written back coalesced.
#pragma once

#include "gp_common.cuh"

namespace gp {

constexpr int kTileThreads = 256;
#ifndef GP_MIN_BLOCKS
#define GP_MIN_BLOCKS 4
#endif
constexpr int kTileMinBlocks = GP_MIN_BLOCKS;   // CTAs per SM the tiled kernels are register-budgeted for

// ---------------------------------------------------------------------------------------------------
// CTA decode + per-CTA sampling records shared by forward and backward
// ---------------------------------------------------------------------------------------------------
struct TileCtx {
    int b, oh0, ow0, g0, TP, n_ul;
};

// tile_h, tile_w and gs are powers of two (plan_tiled): unit / pixel decode is shifts and masks
struct UnitPos {
    int gl, pix, oh, ow;
};
__device__ __forceinline__ UnitPos unit_pos(int ul, const KParams &p, const TileCtx &t) {
    UnitPos u;
    u.gl = ul >> p.lg_tp;
    u.pix = ul & (t.TP - 1);
    u.oh = t.oh0 + (u.pix >> p.lg_tw);
    u.ow = t.ow0 + (u.pix & (p.tile_w - 1));
    return u;
}

__device__ __forceinline__ TileCtx decode_tile(const KParams &p) {
    TileCtx t;
    int bid = blockIdx.x;
    const int gch = bid % p.gchunks;
    bid /= p.gchunks;
    const int tx = bid % p.tiles_x;
    bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    t.b = bid / p.tiles_y;
    t.oh0 = ty * p.tile_h;
    t.ow0 = tx * p.tile_w;
    t.g0 = gch * p.gs;
    t.TP = p.tile_h * p.tile_w;
    t.n_ul = t.TP * p.gs;
    return t;
}

constexpr unsigned F_ALL = 32u;   // backward records: all four corners inside the image (four unpredicated loads / reductions)
constexpr unsigned FW_SX = 1u;    // forward records: both columns of the 2x2 footprint are inside the image
constexpr unsigned FW_SY = 2u;    //                  both rows are
constexpr unsigned FW_OUT = 4u;   // build-time marker of an out-of-range sample (never seen by the sampling loop)

// Sampling records, built ONCE per CTA by ONE thread per unit (its P points in a row, so everything that only depends
// on the unit -- pixel decode, row pointers, p0, the softmax of the fused variant -- is computed once, and for P == 9 the
// kernel indices are compile-time constants) and then read (LDS broadcast) by the L lanes that own the unit's channels.
// The reference redoes this arithmetic in every channel thread (cuh:249-269 run by every thread).  locate() is the same
// device function the index hook runs.  r = ul*P + pt with ul = g_local*TP + pix.
//
//   forward   s_bf[r] = { byte offset of the footprint's first VALID pixel relative to (image, group), FW_SX | FW_SY }
//             s_w [r] = { w1, w2, w3, w4 } * mask, 0 for corners outside the image (cuh:55-76)
//             The four loads of a point go to base, base + sx, base + sy, base + sx + sy with sx / sy = 0 when that side of
//             the footprint is outside the image: every address is a pixel the reference reads for this very sample, the
//             weight of a corner it does not read is 0, so the loop is branch-free for border and interior units alike.
//             A sample that is out of range altogether (cuh:268-269) gets zero weights and borrows the address of another
//             sample of its unit; a unit without any in-range sample is marked dead (s_unit[ul] == 0) and writes zeros.
//   backward  s_bf[r] = { byte offset of corner 1 (may lie outside), F_IN | F_C1..4 | F_ALL },  s_w[r] = { lh, lw, mask, - };
//             after the unit is processed s_w[r] = { grad_off_w, grad_off_h, grad_mask, - }.  s_unit[ul] != 0 <=> all P points
//             have all four corners inside.
// smem: n_rec * 24 bytes + 4 bytes per unit.
template <typename T, bool SOFTMAX, bool BWD, bool P9>
__device__ __forceinline__ void build_records(const T *__restrict__ off, const T *__restrict__ msk, float4 *s_w,
                                              int2 *s_bf, unsigned *s_unit, const KParams &p, const TileCtx &t) {
    const int P = P9 ? 9 : p.P;
    const int cidx = (p.kw / 2) * p.kh + p.kh / 2;   // the centre point in the full kw*kh enumeration
    const int C = p.C, WC = p.W * C;
    // e enumerates (pix, g_local) with g_local fastest: global-memory order of the offset / mask rows
    for (int e = threadIdx.x; e < t.n_ul; e += blockDim.x) {
        const int pix = e >> p.lg_gs, gl = e & (p.gs - 1);
        const int oh = t.oh0 + (pix >> p.lg_tw), ow = t.ow0 + (pix & (p.tile_w - 1));
        const int ul = gl * t.TP + pix;
        if (oh >= p.Ho || ow >= p.Wo) {   // tile overhang: the sampling loop never visits this unit
            s_unit[ul] = 0u;
            continue;
        }
        const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
        // dense rows: (q*G+g)*P, cuh:243-244 (off_q = G*P*2, msk_q = G*P); packed offset||logits rows share one pitch
        const T *orow = off + q * p.off_q + (t.g0 + gl) * (P * 2), *mrow = msk + q * p.msk_q + (t.g0 + gl) * P;
        const float p0_h_ = origin<float>(p.base_h + oh * p.sh, p.half_h, p.scale);
        const float p0_w_ = origin<float>(p.base_w + ow * p.sw, p.half_w, p.scale);
        float mx = 0.f, inv = 1.f;
        if (SOFTMAX) {   // softmax over the P logits of the (pixel, group) row: modules/dcnv3.py:332-333
            mx = to_acc<T>(__ldg(mrow));
            for (int i = 1; i < P; ++i) mx = fmaxf(mx, to_acc<T>(__ldg(mrow + i)));
            float sum = 0.f;
            for (int i = 0; i < P; ++i) sum += expf(to_acc<T>(__ldg(mrow + i)) - mx);
            inv = 1.f / sum;
        }
        float4 *rw = s_w + ul * P;
        int2 *rb = s_bf + ul * P;
        unsigned interior = 1u, out_mask = 0u;
        int ubase = -1;
        auto one_point = [&](int pt, int i, int j) {   // p = i*kh + j, kernel WIDTH index i is the slow one (cuh:257-258)
            float ox, oy;
            load_pair<T>(orow + 2 * pt, ox, oy);   // (w, h) pair, cuh:261-262
            float m = to_acc<T>(__ldg(mrow + pt));
            if (SOFTMAX) m = expf(m - mx) * inv;
            Point<float> sp;
            locate<float>(sp, p0_h_, p0_w_, j * p.dh, i * p.dw, ox, oy, p.scale, p.H, p.W);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            int2 bf = make_int2(0, 0);
            if (BWD) {
                if (sp.flags & F_IN) {
                    bf.x = (sp.h_low * WC + sp.w_low * C) * (int)sizeof(T);
                    bf.y = (int)(sp.flags | ((sp.flags & 30u) == 30u ? F_ALL : 0u));
                    // spare word: the footprint cell (h_low + 1, w_low + 1), read by the in-SM aggregation of dcnv3_bwd_fused
                    w = make_float4(sp.lh, sp.lw, m, __int_as_float(((sp.h_low + 1) << 16) | (sp.w_low + 1)));
                }
                if (!(bf.y & (int)F_ALL)) interior = 0u;
            } else if (sp.flags & F_IN) {
                // cuh:76 weights, mask folded in; corners outside the image contribute 0 (cuh:55-75)
                w.x = (sp.flags & F_C1) ? sp.hh * sp.hw * m : 0.f;
                w.y = (sp.flags & F_C2) ? sp.hh * sp.lw * m : 0.f;
                w.z = (sp.flags & F_C3) ? sp.lh * sp.hw * m : 0.f;
                w.w = (sp.flags & F_C4) ? sp.lh * sp.lw * m : 0.f;
                const bool hl = sp.h_low >= 0, wl = sp.w_low >= 0;
                const bool hh = sp.h_low + 1 <= p.H - 1, wh = sp.w_low + 1 <= p.W - 1;
                const int r0 = hl ? sp.h_low : sp.h_low + 1, c0 = wl ? sp.w_low : sp.w_low + 1;
                bf.x = (r0 * WC + c0 * C) * (int)sizeof(T);
                bf.y = (int)((wl && wh ? FW_SX : 0u) | (hl && hh ? FW_SY : 0u));
                if (ubase < 0) ubase = bf.x;
            } else {
                out_mask |= 1u << (pt & 31);
                bf.y = (int)FW_OUT;   // marker for the borrow pass below (cleared there)
            }
            rw[pt] = w;
            rb[pt] = bf;
        };
        if (P9) {
#pragma unroll
            for (int pt = 0; pt < 9; ++pt) one_point(pt, pt / 3, pt % 3);
        } else {
            for (int pt = 0; pt < P; ++pt) {
                int kk = pt;
                if (p.remove_center && kk >= cidx) ++kk;
                const int i = kk / p.kh;
                one_point(pt, i, kk - i * p.kh);
            }
        }
        if (BWD) {
            s_unit[ul] = interior;
        } else {
            s_unit[ul] = ubase >= 0 ? 1u : 0u;
            if (ubase >= 0 && out_mask) {   // out-of-range samples borrow an address this unit reads anyway (their weights are 0)
                if (P9) {
#pragma unroll
                    for (int pt = 0; pt < 9; ++pt)
                        if (out_mask & (1u << pt)) rb[pt] = make_int2(ubase, 0);
                } else {
                    for (int pt = 0; pt < P; ++pt)
                        if (rb[pt].y == (int)FW_OUT) rb[pt] = make_int2(ubase, 0);
                }
            }
        }
    }
    __syncthreads();
}

// the four corner loads of one sampling point; corners outside the image read nothing and count as 0
// (addresses are byte pointers + 32-bit byte strides: two integer instructions per corner)
template <typename T, int VEC>
__device__ __forceinline__ void gather4(const char *p1, int Cb, int WCb, unsigned flags, float (&v1)[VEC],
                                        float (&v2)[VEC], float (&v3)[VEC], float (&v4)[VEC]) {
    const char *p3 = p1 + WCb;
    if (flags & F_ALL) {
        Vec<T, VEC>::load(reinterpret_cast<const T *>(p1), v1);
        Vec<T, VEC>::load(reinterpret_cast<const T *>(p1 + Cb), v2);
        Vec<T, VEC>::load(reinterpret_cast<const T *>(p3), v3);
        Vec<T, VEC>::load(reinterpret_cast<const T *>(p3 + Cb), v4);
    } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) v1[c] = v2[c] = v3[c] = v4[c] = 0.f;
        if (flags & F_C1) Vec<T, VEC>::load(reinterpret_cast<const T *>(p1), v1);
        if (flags & F_C2) Vec<T, VEC>::load(reinterpret_cast<const T *>(p1 + Cb), v2);
        if (flags & F_C3) Vec<T, VEC>::load(reinterpret_cast<const T *>(p3), v3);
        if (flags & F_C4) Vec<T, VEC>::load(reinterpret_cast<const T *>(p3 + Cb), v4);
    }
}

// ---------------------------------------------------------------------------------------------------
// Forward, tiled + vectorised.  P9 = 9 sampling points (3x3 without remove_center), fully unrolled.
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC, int L, bool P9, bool SOFTMAX>
__global__ void __launch_bounds__(kTileThreads, kTileMinBlocks)
dcnv3_fwd_tile(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk, T *__restrict__ out,
               const __grid_constant__ KParams p) {
    extern __shared__ float4 smem4[];
    const TileCtx t = decode_tile(p);
    const int P = P9 ? 9 : p.P;
    const int n_rec = t.n_ul * P;
    float4 *s_w = smem4;
    int2 *s_bf = reinterpret_cast<int2 *>(s_w + n_rec);
    unsigned *s_unit = reinterpret_cast<unsigned *>(s_bf + n_rec);
    build_records<T, SOFTMAX, false, P9>(off, msk, s_w, s_bf, s_unit, p, t);

    const int cl = threadIdx.x % L;
    const int C = p.C, WC = p.W * C;
    const unsigned Cb = (unsigned)(C * (int)sizeof(T)), WCb = (unsigned)(WC * (int)sizeof(T));
    const T *in_b = in + (long long)t.b * p.H * WC + cl * VEC;
    constexpr int UPB = kTileThreads / L;   // units per pass

    const int n_pass = (t.n_ul + UPB - 1) / UPB;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int ul = pass * UPB + threadIdx.x / L;
        const UnitPos u = unit_pos(ul, p, t);
        if (!(ul < t.n_ul && u.oh < p.Ho && u.ow < p.Wo)) continue;
        const int g = t.g0 + u.gl;
        const char *in_g = reinterpret_cast<const char *>(in_b + g * p.gc);
        const float4 *rw = s_w + ul * P;
        const int2 *rb = s_bf + ul * P;

        float acc[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = 0.f;

        // (w1 v1 + w2 v2 + w3 v3 + w4 v4) * mask, cuh:78 + :270-273 (mask folded into the record's weights); four
        // unconditional gathers per point, no branches -> the scheduler overlaps points freely
        auto point = [&](int k) {
            const int2 bf = rb[k];
            const float4 w = rw[k];
            const unsigned fx = (unsigned)bf.y & FW_SX, fy = (unsigned)bf.y >> 1;
            const char *p1 = in_g + (unsigned)bf.x;
            const char *p2 = p1 + (size_t)fx * Cb;
            const char *p3 = p1 + (size_t)fy * WCb;
            const char *p4 = p3 + (size_t)fx * Cb;
            float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
            Vec<T, VEC>::load(reinterpret_cast<const T *>(p1), v1);
            Vec<T, VEC>::load(reinterpret_cast<const T *>(p2), v2);
            Vec<T, VEC>::load(reinterpret_cast<const T *>(p3), v3);
            Vec<T, VEC>::load(reinterpret_cast<const T *>(p4), v4);
#pragma unroll
            for (int c = 0; c < VEC; ++c)
                acc[c] = fmaf(w.x, v1[c], fmaf(w.y, v2[c], fmaf(w.z, v3[c], fmaf(w.w, v4[c], acc[c]))));
        };
        if (s_unit[ul]) {   // a unit whose samples are all out of range reads nothing and writes zeros (cuh:268-269)
            if (P9) {
#pragma unroll
                for (int k = 0; k < 9; ++k) point(k);
            } else {
                for (int k = 0; k < P; ++k) point(k);
            }
        }
        const long long q = ((long long)t.b * p.Ho + u.oh) * p.Wo + u.ow;
        Vec<T, VEC>::store_stream(out + q * C + g * p.gc + cl * VEC, acc);
    }
}

// ---------------------------------------------------------------------------------------------------
// Forward, generic: any gc, any dtype (incl. f64).  One thread per output scalar.
// ---------------------------------------------------------------------------------------------------
template <typename T, bool SOFTMAX>
__global__ void __launch_bounds__(256)
dcnv3_fwd_generic(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk, T *__restrict__ out,
                  const __grid_constant__ KParams p) {
    using A = typename AccOf<T>::type;
    const long long total = p.n_units * p.gc;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % p.gc);
        const long long unit = idx / p.gc;
        const int g = (int)(unit % p.G);
        const long long q = unit / p.G;
        const int ow = (int)(q % p.Wo), oh = (int)((q / p.Wo) % p.Ho);
        const int b = (int)(q / ((long long)p.Wo * p.Ho));
        const A scale = (A)p.scale;
        const A p0_h_ = origin<A>(p.base_h + oh * p.sh, p.half_h, scale);
        const A p0_w_ = origin<A>(p.base_w + ow * p.sw, p.half_w, scale);
        const long long WC = (long long)p.W * p.C;
        const T *im = in + (long long)b * p.H * WC + g * p.gc + c;
        const T *o = off + unit * (p.P * 2);
        const T *m = msk + unit * p.P;
        A mx = (A)0, inv = (A)1;
        if (SOFTMAX) {
            mx = to_acc<T>(m[0]);
            for (int i = 1; i < p.P; ++i) mx = max(mx, to_acc<T>(m[i]));
            A s = (A)0;
            for (int i = 0; i < p.P; ++i) s += exp(to_acc<T>(m[i]) - mx);
            inv = (A)1 / s;
        }
        const int ch = p.kh / 2, cw = p.kw / 2;
        A col = (A)0;
        int k = 0;
        for (int i = 0; i < p.kw; ++i)
            for (int j = 0; j < p.kh; ++j)
                if (i != cw || j != ch || !p.remove_center) {
                    Point<A> pt;
                    locate<A>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, to_acc<T>(o[2 * k]), to_acc<T>(o[2 * k + 1]), scale,
                              p.H, p.W);
                    A w = to_acc<T>(m[k]);
                    if (SOFTMAX) w = exp(w - mx) * inv;
                    if (pt.flags & F_IN) {
                        const long long base = (long long)pt.h_low * WC + (long long)pt.w_low * p.C;
                        const A v1 = (pt.flags & F_C1) ? to_acc<T>(im[base]) : (A)0;
                        const A v2 = (pt.flags & F_C2) ? to_acc<T>(im[base + p.C]) : (A)0;
                        const A v3 = (pt.flags & F_C3) ? to_acc<T>(im[base + WC]) : (A)0;
                        const A v4 = (pt.flags & F_C4) ? to_acc<T>(im[base + WC + p.C]) : (A)0;
                        const A w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw, w4 = pt.lh * pt.lw;
                        col += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * w;
                    }
                    ++k;
                }
        out[idx] = from_acc<T, A>(col);
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward, tiled + vectorised.  gin accumulates in fp32 (grad_input itself for T=float, the workspace
// for 16-bit storage -- the reference does the same for half, dcnv3_cuda.cu:126-133).
// ---------------------------------------------------------------------------------------------------

// Sum (a, b, c) over the L lanes of a unit with a transposing butterfly: log2(L)+1 shuffles instead of
// 3*log2(L).  On return lane 0 of the unit holds sum(a), lane 2 (L>=4; lane 0 for L==2) sum(b), lane 1 sum(c).
template <int L>
__device__ __forceinline__ void unit_reduce3(float &a, float &b, float &c, int cl) {
    if (L == 1) return;
    const unsigned full = 0xffffffffu;
    const bool odd = cl & 1;
    // xor 1: even lanes keep (a, b), odd lanes keep (c, -)
    const float x = __shfl_xor_sync(full, odd ? a : c, 1);
    const float y = __shfl_xor_sync(full, b, 1);
    float p0 = odd ? c + x : a + x;   // even: a, odd: c
    float p1 = b + y;                 // even: b (odd lanes' copy is unused)
    if (L >= 4) {
        // xor 2: among even lanes bit1==0 keeps a, bit1==1 keeps b; odd lanes keep c
        const bool hi = cl & 2;
        const float z = __shfl_xor_sync(full, (hi || odd) ? p0 : p1, 2);
        p0 = (hi && !odd) ? p1 + z : p0 + z;
#pragma unroll
        for (int o = 4; o < L; o <<= 1) p0 += __shfl_xor_sync(full, p0, o);
        a = b = c = p0;   // lane 0: a, lane 2: b, lane 1: c
    } else {
        a = p0;   // lane 0: a, lane 1: c
        b = p1;   // lane 0: b
        c = p0;
    }
}

// GIN = false: grad_offset / grad_mask only (grad_input comes from dcnv3_gin_binned, dcnv3_gin_binned.cuh)
template <typename T, int VEC, int L, bool P9, bool GIN>
__global__ void __launch_bounds__(kTileThreads, kTileMinBlocks)
dcnv3_bwd_tile(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk,
               const T *__restrict__ gout, float *__restrict__ gin, T *__restrict__ goff, T *__restrict__ gmsk,
               const __grid_constant__ KParams p) {
    extern __shared__ float4 smem4[];
    const TileCtx t = decode_tile(p);
    const int P = P9 ? 9 : p.P;
    const int n_rec = t.n_ul * P;
    float4 *s_w = smem4;
    int2 *s_bf = reinterpret_cast<int2 *>(s_w + n_rec);
    unsigned *s_unit = reinterpret_cast<unsigned *>(s_bf + n_rec);
    build_records<T, false, true, P9>(off, msk, s_w, s_bf, s_unit, p, t);

    const int cl = threadIdx.x % L;
    const int C = p.C, WC = p.W * C;
    const int Cb = C * (int)sizeof(T), WCb = WC * (int)sizeof(T);
    constexpr int GS = (int)(sizeof(float) / sizeof(T));   // byte-offset scale from T storage to the fp32 accumulation image
    const long long img = (long long)t.b * p.H * WC;
    const T *in_b = in + img + cl * VEC;
    float *gin_b = gin + img + cl * VEC;
    constexpr int UPB = kTileThreads / L;
    const int n_pass = (t.n_ul + UPB - 1) / UPB;
    const unsigned full = 0xffffffffu;

    for (int pass = 0; pass < n_pass; ++pass) {
        const int ul = pass * UPB + threadIdx.x / L;
        const UnitPos u = unit_pos(ul, p, t);
        const bool valid = ul < t.n_ul && u.oh < p.Ho && u.ow < p.Wo;
        if (!__any_sync(full, valid)) continue;   // warp-uniform
        const int ulc = valid ? ul : 0;
        const int g = t.g0 + (valid ? u.gl : 0);
        const char *in_g = reinterpret_cast<const char *>(in_b + g * p.gc);
        char *gin_g = reinterpret_cast<char *>(gin_b + g * p.gc);
        float4 *rw = s_w + ulc * P;
        const int2 *rb = s_bf + ulc * P;

        float go[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) go[c] = 0.f;
        if (valid) {
            const long long q = ((long long)t.b * p.Ho + u.oh) * p.Wo + u.ow;
            Vec<T, VEC>::load_stream(gout + q * C + g * p.gc + cl * VEC, go);
        }

        // everything grad_offset / grad_mask need are the four dot products d_k = sum_c top_grad[c] * corner_k[c]
        // (cuh:107-146 are linear in the corner values); s_m / s_w / s_h are this lane's partial sums
        auto point_math = [&](const float4 r, const float (&v1)[VEC], const float (&v2)[VEC], const float (&v3)[VEC],
                              const float (&v4)[VEC], float &s_m, float &s_w_, float &s_h, float (&mw)[4]) {
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
            s_m = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;      // cuh:144  sum_c top_grad * val
            s_w_ = m * (hh * (d2 - d1) + lh * (d4 - d3));     // cuh:145  grad_w_weight * top_grad * mask
            s_h = m * (hw * (d3 - d1) + lw * (d4 - d2));      // cuh:146  grad_h_weight * top_grad * mask
            mw[0] = w1 * m; mw[1] = w2 * m; mw[2] = w3 * m; mw[3] = w4 * m;   // cuh:116-140 corner weights * mask
        };
        // park the unit's three sums in the (consumed) record: lane 0 grad_mask, lane 2 (or 0) grad_off_w, lane 1 grad_off_h
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
            // every unit of the warp is interior: 36 unconditional gathers + 36 unconditional reductions, no branches
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
                float s_m, s_w_, s_h, mw[4];
                point_math(r, v1, v2, v3, v4, s_m, s_w_, s_h, mw);
                char *g1 = gin_g + base * GS, *g3 = g1 + WCb * GS;
#pragma unroll
                for (int c4 = 0; GIN && c4 < VEC; c4 += 4) {
                    red_add_v4(reinterpret_cast<float *>(g1) + c4, mw[0] * go[c4], mw[0] * go[c4 + 1], mw[0] * go[c4 + 2], mw[0] * go[c4 + 3]);
                    red_add_v4(reinterpret_cast<float *>(g1 + Cb * GS) + c4, mw[1] * go[c4], mw[1] * go[c4 + 1], mw[1] * go[c4 + 2], mw[1] * go[c4 + 3]);
                    red_add_v4(reinterpret_cast<float *>(g3) + c4, mw[2] * go[c4], mw[2] * go[c4 + 1], mw[2] * go[c4 + 2], mw[2] * go[c4 + 3]);
                    red_add_v4(reinterpret_cast<float *>(g3 + Cb * GS) + c4, mw[3] * go[c4], mw[3] * go[c4 + 1], mw[3] * go[c4 + 2], mw[3] * go[c4 + 3]);
                }
                unit_reduce3<L>(s_m, s_w_, s_h, cl);
                __syncwarp();   // memory ordering: every lane of the unit has read record k before it is overwritten (shuffles only converge)
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
                    float mw[4];
                    point_math(r, v1, v2, v3, v4, s_m, s_w_, s_h, mw);
                    char *g1 = gin_g + bf.x * GS, *g3 = g1 + WCb * GS;
#pragma unroll
                    for (int c4 = 0; GIN && c4 < VEC; c4 += 4) {
                        if (flags & F_C1) red_add_v4(reinterpret_cast<float *>(g1) + c4, mw[0] * go[c4], mw[0] * go[c4 + 1], mw[0] * go[c4 + 2], mw[0] * go[c4 + 3]);
                        if (flags & F_C2) red_add_v4(reinterpret_cast<float *>(g1 + Cb * GS) + c4, mw[1] * go[c4], mw[1] * go[c4 + 1], mw[1] * go[c4 + 2], mw[1] * go[c4 + 3]);
                        if (flags & F_C3) red_add_v4(reinterpret_cast<float *>(g3) + c4, mw[2] * go[c4], mw[2] * go[c4 + 1], mw[2] * go[c4 + 2], mw[2] * go[c4 + 3]);
                        if (flags & F_C4) red_add_v4(reinterpret_cast<float *>(g3 + Cb * GS) + c4, mw[3] * go[c4], mw[3] * go[c4 + 1], mw[3] * go[c4 + 2], mw[3] * go[c4 + 3]);
                    }
                }
                // sum over the gc channels of the group = reduction over the L lanes of this unit (no block barriers)
                unit_reduce3<L>(s_m, s_w_, s_h, cl);
                __syncwarp();
                if (valid) park(k, s_m, s_w_, s_h);
            }
        }
    }
    __syncthreads();

    // coalesced write-back of grad_offset / grad_mask in global-memory order (inverse of the record permutation)
    for (int e = threadIdx.x; e < n_rec; e += blockDim.x) {
        const int pg = e / P, pt = e - pg * P;
        const int pix = pg >> p.lg_gs, gl = pg & (p.gs - 1);
        const int oh = t.oh0 + (pix >> p.lg_tw), ow = t.ow0 + (pix & (p.tile_w - 1));
        if (oh < p.Ho && ow < p.Wo) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            const long long k = ((q * p.G + t.g0 + gl) * (long long)P) + pt;
            const float4 res = s_w[(gl * t.TP + pix) * P + pt];   // out-of-range samples parked 0, 0, 0 (cuh:347-355)
            store_pair<T>(goff + 2 * k, res.x, res.y);
            gmsk[k] = from_acc<T, float>(res.z);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward, generic: any gc / dtype.  One CTA (64 threads) walks units; threads stride over channels.
// GA = accumulate type of grad_input (float workspace for 16-bit storage).
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(64)
dcnv3_bwd_generic(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk,
                  const T *__restrict__ gout, typename AccOf<T>::type *__restrict__ gin, T *__restrict__ goff,
                  T *__restrict__ gmsk, const __grid_constant__ KParams p) {
    using A = typename AccOf<T>::type;
    __shared__ A red[2][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long WC = (long long)p.W * p.C;
    for (long long unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int g = (int)(unit % p.G);
        const long long q = unit / p.G;
        const int ow = (int)(q % p.Wo), oh = (int)((q / p.Wo) % p.Ho);
        const int b = (int)(q / ((long long)p.Wo * p.Ho));
        const A scale = (A)p.scale;
        const A p0_h_ = origin<A>(p.base_h + oh * p.sh, p.half_h, scale);
        const A p0_w_ = origin<A>(p.base_w + ow * p.sw, p.half_w, scale);
        const long long img = (long long)b * p.H * WC + g * p.gc;
        const T *o = off + unit * (p.P * 2);
        const T *m = msk + unit * p.P;
        const T *go = gout + q * p.C + g * p.gc;
        const int ch = p.kh / 2, cw = p.kw / 2;
        int k = 0;
        for (int i = 0; i < p.kw; ++i)
            for (int j = 0; j < p.kh; ++j)
                if (i != cw || j != ch || !p.remove_center) {
                    Point<A> pt;
                    locate<A>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, to_acc<T>(o[2 * k]), to_acc<T>(o[2 * k + 1]), scale,
                              p.H, p.W);
                    const A w = to_acc<T>(m[k]);
                    A s_m = (A)0, s_w = (A)0, s_h = (A)0;
                    if (pt.flags & F_IN) {
                        const long long base = img + (long long)pt.h_low * WC + (long long)pt.w_low * p.C;
                        const A w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw, w4 = pt.lh * pt.lw;
                        for (int c = threadIdx.x; c < p.gc; c += blockDim.x) {
                            const A tgrad = to_acc<T>(go[c]);
                            const A tg = tgrad * w;
                            A v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                            if (pt.flags & F_C1) { v1 = to_acc<T>(in[base + c]); atomicAdd(gin + base + c, w1 * tg); }
                            if (pt.flags & F_C2) { v2 = to_acc<T>(in[base + p.C + c]); atomicAdd(gin + base + p.C + c, w2 * tg); }
                            if (pt.flags & F_C3) { v3 = to_acc<T>(in[base + WC + c]); atomicAdd(gin + base + WC + c, w3 * tg); }
                            if (pt.flags & F_C4) { v4 = to_acc<T>(in[base + WC + p.C + c]); atomicAdd(gin + base + WC + p.C + c, w4 * tg); }
                            const A val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                            const A gw = pt.hh * (v2 - v1) + pt.lh * (v4 - v3);
                            const A gh = pt.hw * (v3 - v1) + pt.lw * (v4 - v2);
                            s_m += tgrad * val;
                            s_w += gw * tg;
                            s_h += gh * tg;
                        }
                    }
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) {
                        s_m += __shfl_xor_sync(0xffffffffu, s_m, o2);
                        s_w += __shfl_xor_sync(0xffffffffu, s_w, o2);
                        s_h += __shfl_xor_sync(0xffffffffu, s_h, o2);
                    }
                    if (lane == 0) { red[warp][0] = s_m; red[warp][1] = s_w; red[warp][2] = s_h; }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        gmsk[unit * p.P + k] = from_acc<T, A>(red[0][0] + red[1][0]);
                        goff[(unit * p.P + k) * 2] = from_acc<T, A>(scale * (red[0][1] + red[1][1]));
                        goff[(unit * p.P + k) * 2 + 1] = from_acc<T, A>(scale * (red[0][2] + red[1][2]));
                    }
                    __syncthreads();
                    ++k;
                }
    }
}

// fp32 accumulation image -> 16-bit grad_input (dcnv3_cuda.cu:168-173)
template <typename T>
__global__ void __launch_bounds__(256) cast_from_f32(const float *__restrict__ src, T *__restrict__ dst, long long n) {
    const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n) {
        const float4 a = *reinterpret_cast<const float4 *>(src + i), b = *reinterpret_cast<const float4 *>(src + i + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        Vec<T, 8>::store_stream(dst + i, v);
    } else {
        for (long long k = i; k < n; ++k) dst[k] = from_acc<T, float>(src[k]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Index hook (parity of floor()/bounds): one thread per unit.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
dcnv3_index_kernel(const T *__restrict__ off, int *__restrict__ hw_low, unsigned char *__restrict__ flags,
                   const __grid_constant__ KParams p) {
    using A = typename AccOf<T>::type;
    const long long unit = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (unit >= p.n_units) return;
    const long long q = unit / p.G;
    const int ow = (int)(q % p.Wo), oh = (int)((q / p.Wo) % p.Ho);
    const A scale = (A)p.scale;
    const A p0_h_ = origin<A>(p.base_h + oh * p.sh, p.half_h, scale);
    const A p0_w_ = origin<A>(p.base_w + ow * p.sw, p.half_w, scale);
    const int ch = p.kh / 2, cw = p.kw / 2;
    long long k = unit * p.P;
    for (int i = 0; i < p.kw; ++i)
        for (int j = 0; j < p.kh; ++j)
            if (i != cw || j != ch || !p.remove_center) {
                Point<A> pt;
                locate<A>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, to_acc<T>(off[2 * k]), to_acc<T>(off[2 * k + 1]), scale,
                          p.H, p.W);
                hw_low[2 * k] = pt.h_low;
                hw_low[2 * k + 1] = pt.w_low;
                flags[k] = (unsigned char)pt.flags;
                ++k;
            }
}

}  // namespace gp
