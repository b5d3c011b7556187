// givepose_b200 -- shared device helpers for the DCNv3 kernels (sm_100a).
//
// The sampling geometry below is the one definition used by the forward kernels, the backward kernels
// and the gp_dcnv3_sample_index parity hook, so "indices and bounds are bit-exact" is checked on the
// code that actually runs.  Arithmetic follows the reference kernel
// (network/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:232-269 for the location, :39-75 for the corners)
// in its accumulate type (fp32 for f32/bf16/f16 storage, fp64 for f64).  The two products that nvcc
// contracts into FMAs in the reference build (p0_ - c*scale, p0_ + (i*dil+off)*scale) are written as
// explicit fma() so the rounding does not depend on compiler flags.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "givepose_b200.h"

namespace gp {

// ---- kernel-side copy of the call geometry ---------------------------------------------------------
struct KParams {
    int N, H, W, G, gc, C;
    int kh, kw, sh, sw, ph, pw, dh, dw;
    int remove_center, P;
    int Ho, Wo;
    int base_h, base_w;      // ((dil*(k-1))>>1) - pad                     (cuh:232-236 without the ow*stride term)
    int half_h, half_w;      // (dil*(k-1))>>1
    float scale;
    // tiling of the output plane used by the tiled kernels
    int tile_h, tile_w, tiles_y, tiles_x, gs /*groups per CTA*/, gchunks /*G/gs*/;   // tile_h, tile_w, gs: powers of two
    int lg_tw, lg_tp, lg_gs; // log2(tile_w), log2(tile_h*tile_w), log2(gs)
    long long n_units;       // N*Ho*Wo*G
    // offset / mask row addressing of the tiled forward: offset row of (q, g) starts at off + q*off_q + g*P*2, mask row at
    // msk + q*msk_q + g*P.  Dense tensors (the reference layout, cuh:243-244): off_q = G*P*2, msk_q = G*P.  The packed
    // offset||mask-logits rows of gp_dcnv3_forward_softmax_packed use one pitch for both.
    long long off_q, msk_q;
};

template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

__device__ __forceinline__ float gp_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double gp_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float gp_floor(float a) { return floorf(a); }
__device__ __forceinline__ double gp_floor(double a) { return floor(a); }

template <typename T> __device__ __forceinline__ typename AccOf<T>::type to_acc(T v);
template <> __device__ __forceinline__ float to_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ double to_acc<double>(double v) { return v; }
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }

template <typename T, typename A> __device__ __forceinline__ T from_acc(A v);
template <> __device__ __forceinline__ float from_acc<float, float>(float v) { return v; }
template <> __device__ __forceinline__ double from_acc<double, double>(double v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }

// ---- one sampling point ------------------------------------------------------------------------------
enum : unsigned { F_IN = 1u, F_C1 = 2u, F_C2 = 4u, F_C3 = 8u, F_C4 = 16u };

template <typename A> struct Point {
    A lh, lw, hh, hw;
    int h_low, w_low;
    unsigned flags;   // F_IN | per-corner validity
};

// p0_*_ of cuh:249-252:  p0 - half*scale   (p0 = base + o*stride)
template <typename A> __device__ __forceinline__ A origin(int p0, int half, A scale) {
    return gp_fma(-(A)half, scale, (A)p0);
}

// cuh:263-269 + :39-75.  idil = i*dilation (integer), off = offset value already widened to A.
template <typename A>
__device__ __forceinline__ void locate(Point<A> &pt, A p0_h_, A p0_w_, int jdil_h, int idil_w, A off_w, A off_h,
                                       A scale, int H, int W) {
    const A loc_w = gp_fma((A)idil_w + off_w, scale, p0_w_);
    const A loc_h = gp_fma((A)jdil_h + off_h, scale, p0_h_);
    pt.flags = 0u;
    pt.h_low = 0;
    pt.w_low = 0;
    pt.lh = pt.lw = pt.hh = pt.hw = (A)0;
    if (loc_h > (A)-1 && loc_w > (A)-1 && loc_h < (A)H && loc_w < (A)W) {
        const A fh = gp_floor(loc_h), fw = gp_floor(loc_w);
        const int h_low = (int)fh, w_low = (int)fw;
        pt.h_low = h_low;
        pt.w_low = w_low;
        pt.lh = loc_h - fh;      // == loc_h - (A)h_low  (fh is integral and exactly representable)
        pt.lw = loc_w - fw;
        pt.hh = (A)1 - pt.lh;
        pt.hw = (A)1 - pt.lw;
        const bool hl = h_low >= 0, wl = w_low >= 0, hh_ok = h_low + 1 <= H - 1, wh_ok = w_low + 1 <= W - 1;
        pt.flags = F_IN | (hl && wl ? F_C1 : 0u) | (hl && wh_ok ? F_C2 : 0u) | (hh_ok && wl ? F_C3 : 0u) |
                   (hh_ok && wh_ok ? F_C4 : 0u);
    }
}

// ---- vector load/store of VEC channels into accumulate-type registers -----------------------------------
template <typename T, int VEC> struct Vec;   // VEC channels of storage type T, 16 bytes when VEC*sizeof(T)==16

template <> struct Vec<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
    static __device__ __forceinline__ void load_stream(const float *p, float (&v)[4]) {
        float4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
    static __device__ __forceinline__ void store_stream(float *p, const float (&v)[4]) {
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
};

template <> struct Vec<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void unpack(const uint4 r, float (&v)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // bf16 -> f32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
        unpack(__ldg(reinterpret_cast<const uint4 *>(p)), v);
    }
    static __device__ __forceinline__ void load_stream(const __nv_bfloat16 *p, float (&v)[8]) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        unpack(r, v);
    }
    static __device__ __forceinline__ void store_stream(__nv_bfloat16 *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
};

template <> struct Vec<__half, 8> {
    static __device__ __forceinline__ void unpack(const uint4 r, float (&v)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void load(const __half *p, float (&v)[8]) {
        unpack(__ldg(reinterpret_cast<const uint4 *>(p)), v);
    }
    static __device__ __forceinline__ void load_stream(const __half *p, float (&v)[8]) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        unpack(r, v);
    }
    static __device__ __forceinline__ void store_stream(__half *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
};

// 16-bit storage, 4 channels per lane (8-byte requests): same L1 wavefront count per unit as the 8-channel
// variant, half the registers / instructions per lane
template <> struct Vec<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void unpack(const uint2 r, float (&v)[4]) {
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[4]) {
        unpack(__ldg(reinterpret_cast<const uint2 *>(p)), v);
    }
    static __device__ __forceinline__ void load_stream(const __nv_bfloat16 *p, float (&v)[4]) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        unpack(r, v);
    }
    static __device__ __forceinline__ void store_stream(__nv_bfloat16 *p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(*reinterpret_cast<const uint32_t *>(&a)),
                     "r"(*reinterpret_cast<const uint32_t *>(&b)) : "memory");
    }
};

template <> struct Vec<__half, 4> {
    static __device__ __forceinline__ void unpack(const uint2 r, float (&v)[4]) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void load(const __half *p, float (&v)[4]) {
        unpack(__ldg(reinterpret_cast<const uint2 *>(p)), v);
    }
    static __device__ __forceinline__ void load_stream(const __half *p, float (&v)[4]) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        unpack(r, v);
    }
    static __device__ __forceinline__ void store_stream(__half *p, const float (&v)[4]) {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(*reinterpret_cast<const uint32_t *>(&a)),
                     "r"(*reinterpret_cast<const uint32_t *>(&b)) : "memory");
    }
};

// (x, y) pair of consecutive elements, e.g. one sampling point's (w, h) offset
template <typename T> __device__ __forceinline__ void load_pair(const T *p, float &a, float &b);
template <> __device__ __forceinline__ void load_pair<float>(const float *p, float &a, float &b) {
    const float2 r = __ldg(reinterpret_cast<const float2 *>(p));
    a = r.x; b = r.y;
}
template <> __device__ __forceinline__ void load_pair<__nv_bfloat16>(const __nv_bfloat16 *p, float &a, float &b) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t *>(p));
    a = __uint_as_float(r << 16); b = __uint_as_float(r & 0xffff0000u);
}
template <> __device__ __forceinline__ void load_pair<__half>(const __half *p, float &a, float &b) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t *>(p));
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&r));
    a = f.x; b = f.y;
}
template <typename T> __device__ __forceinline__ void store_pair(T *p, float a, float b);
template <> __device__ __forceinline__ void store_pair<float>(float *p, float a, float b) {
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void store_pair<__nv_bfloat16>(__nv_bfloat16 *p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b);
}
template <> __device__ __forceinline__ void store_pair<__half>(__half *p, float a, float b) {
    *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(a, b);
}

// fp32 vector reduction into global memory (REDG.E.ADD.F32x4 on sm_90+): one 16-byte atomic per lane
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// launch accounting (gp_launch_count)
extern unsigned long long g_launches;
inline void count_launch(unsigned long long n = 1) { g_launches += n; }

}  // namespace gp
