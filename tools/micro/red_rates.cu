// Micro-benchmark (sm_100a): how fast can one SM push 128-byte fp32 reductions into L2?
//   mode 0  red.global.add.v4.f32, 8 lanes per 128-byte line (what dcnv3_bwd_tile issues)
//   mode 1  red.global.add.f32, 32 lanes per line
//   mode 2  cp.reduce.async.bulk.global.shared::cta.add.f32, one 128-byte bulk op per line (4 lanes of a warp issue)
//   mode 3  as 2 but 512-byte ops (4 consecutive lines)
//   mode 4  shared-memory integer atomicAdd with return value, spread addresses (the ranking step of a binned backward)
//   mode 5  STS.128 of the contribution + mode 2 (what a TMA-reduce backward would do per corner)
// Addresses: every CTA walks pseudo-randomly inside a window of WIN lines that moves every 64 iterations (like the
// sampling windows of a tile), so reductions mostly hit L2.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red_v4(float *a, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void bulk_red(float *g, uint32_t s, int bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float *buf, long long n_lines, int iters, int win, unsigned *sink) {
    __shared__ __align__(128) float stage[8][4][128];   // per warp: 4 lines (mode 3: one 512-byte op)
    __shared__ unsigned cnt[2048];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2048; i += 256) cnt[i] = 0;
    for (int i = lane; i < 512; i += 32) (&stage[warp][0][0])[i] = 1.0f;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t wbase = hash32(blockIdx.x * 7919u + (it >> 6)) % (uint32_t)(n_lines - win);
        if (MODE == 0) {
            const uint32_t line = wbase + hash32(it * 131u + warp * 17u + (lane >> 3) + blockIdx.x) % win;
            red_v4(buf + (long long)line * 32 + (lane & 7) * 4, 1.f, 1.f, 1.f, 1.f);
        } else if (MODE == 1) {
            const uint32_t line = wbase + hash32(it * 131u + warp * 17u + blockIdx.x) % win;
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(buf + (long long)line * 32 + lane), "f"(1.f) : "memory");
        } else if (MODE == 2 || MODE == 5) {
            if (MODE == 5) *reinterpret_cast<float4 *>(&stage[warp][lane >> 3][(lane & 7) * 4]) = make_float4(1.f, 1.f, 1.f, 1.f);
            if (MODE == 5) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); __syncwarp(); }
            if (lane < 4) {
                const uint32_t line = wbase + hash32(it * 131u + warp * 17u + lane + blockIdx.x) % win;
                bulk_red(buf + (long long)line * 32, (uint32_t)__cvta_generic_to_shared(&stage[warp][lane][0]), 128);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if ((it & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (MODE == 5) __syncwarp();
        } else if (MODE == 3) {
            if (lane == 0) {
                const uint32_t line = wbase + hash32(it * 131u + warp * 17u + blockIdx.x) % (win - 4);
                bulk_red(buf + (long long)line * 32, (uint32_t)__cvta_generic_to_shared(&stage[warp][0][0]), 512);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if ((it & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else if (MODE == 4) {
            acc += atomicAdd(&cnt[hash32(it * 131u + threadIdx.x * 2654435761u) & 2047], 1u);
        }
    }
    if (MODE == 2 || MODE == 3 || MODE == 5) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 0xdeadbeef) *sink = acc;
}

template <int MODE> void run(const char *name, float *buf, long long n_lines, int iters, int win, unsigned *sink, double lines_per_warp_iter) {
    const int grid = 148 * 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 256>>>(buf, n_lines, iters / 4, win, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(buf, n_lines, iters, win, sink);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double lines = (double)grid * 8 * iters * lines_per_warp_iter;
    printf("%-58s win %5d  %8.3f ms  %7.2f Glines/s  %6.0f GB/s  %.3f lines/clk/SM @1.9GHz  %s\n", name, win, ms, lines / ms / 1e6,
           lines * 128 / ms / 1e6, lines / (ms * 1e-3) / 148 / 1.9e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const long long n_lines = 268435456LL / 128;   // 268 MB fp32 image batch, like grad_input at config 2
    float *buf; unsigned *sink;
    cudaMalloc(&buf, n_lines * 128); cudaMemset(buf, 0, n_lines * 128); cudaMalloc(&sink, 4);
    const int iters = 4096;
    for (int win : {512, 4096}) {
        run<0>("0 red.v4.f32 (8 lanes/line, 4 lines/warp-instr)", buf, n_lines, iters, win, sink, 4);
        run<1>("1 red.f32 (32 lanes/line)", buf, n_lines, iters, win, sink, 1);
        run<2>("2 cp.reduce.async.bulk 128 B (4 ops/warp-iter)", buf, n_lines, iters, win, sink, 4);
        run<3>("3 cp.reduce.async.bulk 512 B (1 op/warp-iter)", buf, n_lines, iters, win, sink, 4);
        run<5>("5 STS.128 + fence + cp.reduce.async.bulk 128 B", buf, n_lines, iters, win, sink, 4);
    }
    run<4>("4 ATOMS.ADD u32 with return, spread (per LANE-op, not line)", buf, n_lines, iters, 512, sink, 32);
    return 0;
}
