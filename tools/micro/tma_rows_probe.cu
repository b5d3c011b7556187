// Probe: which 3-D non-swizzled TMA box shapes does sm_100a accept?  usage: tma_rows_probe <inner_elems> <box_inner> <box_w> <box_h> <rank>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int n, int rank) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
    dst = (dst + 127u) & ~127u;
    uint32_t bar = dst + ((n * 4 + 127) / 128) * 128;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n * 4) : "memory");
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(dst), "l"(&map), "r"(bar), "r"(0), "r"(0), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(dst), "l"(&map), "r"(bar), "r"(0), "r"(0) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar) : "memory");
    const float *s = reinterpret_cast<const float *>(sm + (dst - (uint32_t)__cvta_generic_to_shared(sm)));
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char **argv) {
    const int inner = atoi(argv[1]), bi = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]), rank = atoi(argv[5]);
    const int W = 64, R = 256;
    float *g, *out;
    cudaMalloc(&g, (size_t)inner * W * R * 4);
    cudaMemset(g, 0, (size_t)inner * W * R * 4);
    const int n = bi * bw * (rank == 3 ? bh : 1);
    cudaMalloc(&out, n * 4);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)W, (cuuint64_t)R};
    cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * 4 * W};
    cuuint32_t box[3] = {(cuuint32_t)bi, (cuuint32_t)bw, (cuuint32_t)bh};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("inner %d box {%d,%d,%d} rank %d: encode rc %d; ", inner, bi, bw, bh, rank, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
    probe<<<1, 128, n * 4 + 512>>>(map, out, n, rank);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    return 0;
}
