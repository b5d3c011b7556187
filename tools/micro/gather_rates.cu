// Micro-benchmark (sm_100a): the speed of light of the DCNv3 forward's access pattern, measured instead of quoted.
//
// One unit = one (output pixel, group) of BASELINE config 2 (N = 64, 64x64x256 channel-last, G = 8, gc = 32, 3x3): 36 corner
// gathers of gc channels + 1 output line.  The kernels below do EXACTLY that and nothing else: the 36 byte offsets per
// unit are precomputed (what a free record builder would hand over), there is no offset / mask traffic, no coordinate
// arithmetic, no bounds handling.  Offsets are drawn like the reference's test distribution (offsets U[0,10) px around
// the 3x3 grid, ops_dcnv3/test.py:36-40) or model-like (N(0,1) px), clipped to the image, so L1 / L2 locality is the real
// kernel's.  CTA = 8x8 pixels x 2 groups, 256 threads, like dcnv3_fwd_tile.
//
//   mode 0  fp32 rows: 8 lanes x 16 B per corner (one full 128-byte line per gather)
//   mode 1  16-bit rows, 8 lanes x 8 B per corner (half a line per gather)
//   mode 2  16-bit rows, 4 lanes x 16 B per corner (half a line per gather, 8 units per warp request)
//   mode 3  mode 0 without the record reads (offsets derived from a per-unit seed in registers: pure gather + store)
//   mode 4  shared-memory reads only: the LDS patterns the kernels use (16-byte broadcast records, 128-byte rows)
//   mode 5  the backward's scatter alone: 36 red.global.add.v4.f32 line reductions per unit to the same addresses (fp32)
//   mode 6  gather + scatter: 36 line gathers AND 36 line reductions per unit (the one-pass backward's memory pattern)
//   mode 7  scatter, best locality: the 36 reductions of a unit all go to ONE line (its own output pixel)
//   mode 8  scatter, no locality: 36 reductions per unit to pseudo-random lines of the whole 268 MB image tensor
//           (7 and 8 bracket mode 5: does the chip's reduction rate depend on WHERE the lines are?)
// Output: ms per launch, units/clk/SM, and for modes 0-3 the bytes-gathered rate.  The forward can not be faster than
// mode 0 (fp32) / mode 1-2 (bf16) with the same decomposition; the figure bench.py quotes as `l1_gather_floor_ms`.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/gather_rates tools/micro/gather_rates.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int N = 64, H = 64, W = 64, G = 8, GC = 32, C = G * GC, P = 9;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ float u01(uint32_t h) { return (h >> 8) * (1.0f / 16777216.0f); }
// approx N(0,1): sum of 4 uniforms, scaled
__device__ __forceinline__ float nrm(uint32_t h) {
    return (u01(hash32(h)) + u01(hash32(h + 1)) + u01(hash32(h + 2)) + u01(hash32(h + 3)) - 2.0f) * 1.7320508f;
}

// element offsets (in PIXELS of the (image, group) plane: y*W + x) of the 4 corners of the 9 points of every unit
__global__ void make_offsets(int *__restrict__ offs, int dist) {
    const long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // unit = ((b*H+oh)*W+ow)*G+g
    if (u >= (long long)N * H * W * G) return;
    const int ow = (int)((u / G) % W), oh = (int)((u / G / W) % H);
    for (int pt = 0; pt < P; ++pt) {
        const uint32_t s = (uint32_t)(u * 31 + pt) * 2654435761u;
        const float dx = dist == 0 ? u01(hash32(s)) * 10.f : nrm(s);
        const float dy = dist == 0 ? u01(hash32(s ^ 0x9e3779b9u)) * 10.f : nrm(s ^ 0x9e3779b9u);
        int x = (int)floorf(ow + pt / 3 - 1 + dx), y = (int)floorf(oh + pt % 3 - 1 + dy);
        x = min(max(x, 0), W - 2); y = min(max(y, 0), H - 2);
        int4 o = make_int4(y * W + x, y * W + x + 1, (y + 1) * W + x, (y + 1) * W + x + 1);
        reinterpret_cast<int4 *>(offs)[u * P + pt] = o;
    }
}

__device__ __forceinline__ void red_v4(float *a, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 4)
gather(const void *__restrict__ in, const int *__restrict__ offs, void *__restrict__ out) {
    constexpr int L = MODE == 2 ? 4 : 8;            // lanes per unit
    constexpr int EB = MODE == 1 || MODE == 2 ? 2 : 4;   // element bytes
    constexpr int LB = GC * EB / L;                 // bytes per lane per corner: 16, 8, 16
    constexpr int UPB = 256 / L;
    // tile decode as dcnv3_fwd_tile: gch fastest, then tile x, tile y, image
    int bid = blockIdx.x;
    const int gch = bid % (G / 2); bid /= (G / 2);
    const int tx = bid % (W / 8); bid /= (W / 8);
    const int ty = bid % (H / 8);
    const int b = bid / (H / 8);
    const int cl = threadIdx.x % L;
    const char *in_b = (const char *)in + (long long)b * H * W * C * EB;
    for (int pass = 0; pass < 128 / UPB; ++pass) {
        const int ul = pass * UPB + threadIdx.x / L;
        const int gl = ul >> 6, pix = ul & 63;
        const int oh = ty * 8 + (pix >> 3), ow = tx * 8 + (pix & 7), g = gch * 2 + gl;
        const long long q = ((long long)b * H + oh) * W + ow;
        const long long u = q * G + g;
        const char *in_g = in_b + (g * GC) * EB + cl * LB;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int pt = 0; pt < P; ++pt) {
            int4 o;
            if (MODE == 3) {
                const uint32_t s = hash32((uint32_t)u * 9u + pt);
                int x = min(max(ow + pt / 3 - 1 + (int)(s & 7), 0), W - 2), y = min(max(oh + pt % 3 - 1 + (int)((s >> 3) & 7), 0), H - 2);
                o = make_int4(y * W + x, y * W + x + 1, (y + 1) * W + x, (y + 1) * W + x + 1);
            } else {
                o = __ldg(reinterpret_cast<const int4 *>(offs) + u * P + pt);
            }
            const int oo[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const char *ptr = in_g + (long long)oo[k] * (C * EB);
                if (MODE == 7) {
                    red_v4(reinterpret_cast<float *>((char *)out + (q * C + g * GC) * 4 + cl * 16), 1.f, 1.f, 1.f, 1.f);
                } else if (MODE == 8) {
                    const uint32_t line = hash32((uint32_t)u * 36u + pt * 4 + k) % (uint32_t)(N * H * W * G);
                    red_v4(reinterpret_cast<float *>((char *)out + (size_t)line * 128 + cl * 16), 1.f, 1.f, 1.f, 1.f);
                } else if (MODE == 5 || MODE == 6) {   // `out` is the fp32 accumulation image (same shape as `in`)
                    float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (MODE == 6) v = __ldg(reinterpret_cast<const float4 *>(ptr));
                    red_v4(reinterpret_cast<float *>((char *)out + (ptr - (const char *)in)), v.x, v.y, v.z, v.w);
                } else if (LB == 16) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(ptr));
                    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
                } else {
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(ptr));
                    acc[0] += v.x; acc[1] += v.y;
                }
            }
        }
        char *o_ptr = (char *)out + (q * C + g * GC) * EB + cl * LB;
        if (MODE >= 5) continue;
        if (LB == 16) *reinterpret_cast<float4 *>(o_ptr) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else *reinterpret_cast<float2 *>(o_ptr) = make_float2(acc[0], acc[1]);
    }
}

// mode 4: LDS patterns.  pattern 0: LDS.128, 4 distinct 16-byte records per warp (8 lanes broadcast each)
//                         pattern 1: LDS.128, 4 rows of 128 bytes per warp (8 lanes x 16 B each), random rows
//                         pattern 2: LDS.64 broadcast records (4 distinct per warp)
//                         pattern 3: LDS.128, 8 half rows (4 lanes x 16 B = 64 B each), random rows (bf16 window reads)
//                         pattern 4: LDS.128 all 32 lanes the same address
template <int PAT>
__global__ void __launch_bounds__(256, 4) lds_rates(float *sink, int iters) {
    __shared__ __align__(16) float4 s[2048];   // 32 KB
    for (int i = threadIdx.x; i < 2048; i += 256) s[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a = 0.f;
    uint32_t h = hash32(threadIdx.x / (PAT == 3 ? 4 : 8) + 977u * blockIdx.x);
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        if (PAT == 0) { const float4 v = s[(h >> 8) & 2047]; a += v.x + v.w; }
        else if (PAT == 1) { const float4 v = s[(((h >> 8) & 255) << 3) + (lane & 7)]; a += v.x + v.w; }
        else if (PAT == 2) { const float2 v = reinterpret_cast<const float2 *>(s)[(h >> 8) & 4095]; a += v.x + v.y; }
        else if (PAT == 3) { const float4 v = s[(((h >> 8) & 511) << 2) + (lane & 3)]; a += v.x + v.w; }
        else { const float4 v = s[(it * 7 + warp) & 2047]; a += v.x + v.w; }
    }
    if (a == 123.456f) *sink = a;
}

template <int MODE> void run(const char *name, const void *in, const int *offs, void *out, int dist) {
    const int grid = N * (H / 8) * (W / 8) * (G / 2);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) gather<MODE><<<grid, 256>>>(in, offs, out);
    cudaDeviceSynchronize();
    const int reps = 10;
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) gather<MODE><<<grid, 256>>>(in, offs, out);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    float ms = 0; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    const double units = (double)N * H * W * G;
    const int eb = (MODE == 1 || MODE == 2) ? 2 : 4;
    printf("gather mode %d dist %c %-44s %7.4f ms  %6.2f clk/unit/SM @1.965GHz  gathered %6.0f GB/s  %s\n", MODE, dist ? 'M' : 'T', name,
           ms, ms * 1e-3 * 1.965e9 / (units / 148), units * 36 * GC * eb / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int PAT> void run_lds(const char *name, float *sink) {
    const int grid = 148 * 4, iters = 8192;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    lds_rates<PAT><<<grid, 256>>>(sink, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    lds_rates<PAT><<<grid, 256>>>(sink, iters);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double winstr = (double)4 * 8 * iters;   // warp instructions per SM
    printf("lds pattern %d %-64s %7.4f ms  %5.2f clk per warp-LDS per SM @1.965GHz\n", PAT, name, ms, ms * 1e-3 * 1.965e9 / winstr);
}

int main() {
    void *in, *out; int *offs; float *sink;
    const size_t n_in = (size_t)N * H * W * C * 4;
    cudaMalloc(&in, n_in); cudaMalloc(&out, n_in); cudaMemset(in, 0, n_in);
    cudaMalloc(&offs, (size_t)N * H * W * G * P * 16); cudaMalloc(&sink, 4);
    for (int dist = 0; dist < 2; ++dist) {
        make_offsets<<<(N * H * W * G + 255) / 256, 256>>>(offs, dist);
        cudaDeviceSynchronize();
        run<0>("fp32, 8 lanes x 16 B (full lines)", in, offs, out, dist);
        run<1>("16-bit, 8 lanes x 8 B (half lines)", in, offs, out, dist);
        run<2>("16-bit, 4 lanes x 16 B (half lines)", in, offs, out, dist);
        run<5>("fp32 scatter only: 36 red.v4 lines per unit", in, offs, out, dist);
        run<6>("fp32 gather + scatter: 36 + 36 lines per unit", in, offs, out, dist);
        if (dist == 0) {
            run<7>("fp32 scatter, 36 reductions to ONE line per unit", in, offs, out, dist);
            run<8>("fp32 scatter, 36 reductions to random lines", in, offs, out, dist);
        }
    }
    run<3>("fp32, offsets from registers (no record reads)", in, offs, out, 0);
    run_lds<0>("LDS.128, 4 distinct 16 B records per warp (8-lane broadcast)", sink);
    run_lds<1>("LDS.128, 4 random 128 B rows per warp", sink);
    run_lds<2>("LDS.64, 4 distinct 8 B records per warp (8-lane broadcast)", sink);
    run_lds<3>("LDS.128, 8 random 64 B half rows per warp", sink);
    run_lds<4>("LDS.128, one address per warp", sink);
    return 0;
}
