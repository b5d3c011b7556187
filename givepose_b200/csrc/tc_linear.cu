// givepose_b200 -- dense layers of the PoseNet heads on the 5th-generation tensor cores (sm_100a):
//
//     y[M,N] = act(x[M,K] . w[N,K]^T + bias[N])          bf16 operands, fp32 accumulation in TMEM, bf16 result
//
// This is the shape of every Linear / 1x1 convolution on the path -- DCNv3's input_proj / output_proj / offset / mask
// (modules/dcnv3.py:325-354, K = 256), feat_reducer (PoseNet.py:158), and the PnP regression trunk fc1||fc1_z, fc2, fc2_z
// (conv_pnp_net.py:172-199, K = 8192 / 1024) with its LeakyReLU(0.1) -- with the bias and the activation applied in the
// epilogue instead of in separate elementwise passes.
//
// Structure: persistent CTAs (one per SM, 192 threads, warp-specialised) walking the 128 x BN output tiles, n fastest so the
// CTAs that share a slab of x run together (the second read hits L2):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d loads of a 128 x 64 slab of x and a BN x 64 slab of w per K step into a
//              ring of 128B-swizzled shared-memory stages, completion signalled on mbarriers (expect_tx); runs ahead across
//              tile boundaries
//   warp 1     allocates 2 x BN TMEM columns (two accumulators); one elected thread issues
//              tcgen05.mma.cta_group::1.kind::f16 (M128 x BN x K16, both operands K-major from shared memory through 64-bit
//              matrix descriptors); tcgen05.commit releases each stage back to the producer and hands the finished
//              accumulator to the epilogue
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns at a time (warp w owns TMEM lanes 32*(w%4)..+31 = output rows),
//              + bias, activation, pack to bf16 into a swizzled staging tile, release the accumulator (so the MMAs of the tile
//              after next start while this one is written out), then row-contiguous 16-byte global stores
// Rows / columns beyond M / N and the K tail are zero-filled by TMA on the way in and masked on the way out.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace gp {

namespace tc {

template <int BN> struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN == 128 ? 4 : 3;
    static constexpr int OUT_BYTES = BM * BN * 2;              // staging tile of the epilogue (32 rows x BN per warp)
    static constexpr int TMEM_COLS = 2 * BN;                   // two fp32 accumulators of 128 lanes x BN columns
    static constexpr int BIAS_BYTES = 4 * BN * 4;               // one fp32 copy of the tile's bias slice per epilogue warp
    static constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)STAGES * STAGE_BYTES + OUT_BYTES + BIAS_BYTES + 256 /*barriers*/;
    // instruction descriptor, kind::f16: D fp32 (bit 4), A/B bf16 (bits 7, 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};


template <int ACT> __device__ __forceinline__ float activate(float v, float slope) {
    if (ACT == ACT_LRELU) return fmaxf(v, v * slope);   // 0 <= slope <= 1 (checked on the host)
    if (ACT == ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

template <int BN, int ACT>
__global__ void __launch_bounds__(THREADS, 1)
linear_bf16_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const float *__restrict__ bias, __nv_bfloat16 *__restrict__ y, int M, int N, int K, float slope) {
    using C = Cfg<BN>;
    constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
    const uint32_t out_stage = base + STAGES * STAGE_BYTES;
    const uint32_t bias_stage = out_stage + C::OUT_BYTES;
    const uint32_t bars = bias_stage + C::BIAS_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);         // 4 bytes: TMEM base address written by tcgen05.alloc
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));      // generic pointer to the aligned base

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blocks = (N + BN - 1) / BN, m_blocks = (M + BM - 1) / BM;
    const long long tiles = (long long)n_blocks * m_blocks;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 4);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {   // one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;   // K steps issued so far, across tiles
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int m0 = (int)(tile / n_blocks) * BM, n0 = (int)(tile % n_blocks) * BN;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);               // slot free (passes immediately on the first lap)
                    mbar_expect_tx(full_bar(s), STAGE_BYTES);
                    tma_load_2d(base + s * STAGE_BYTES, &map_x, full_bar(s), kb * BK, m0);
                    tma_load_2d(base + s * STAGE_BYTES + A_BYTES, &map_w, full_bar(s), kb * BK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t it = 0, lt = 0;   // K steps / tiles consumed so far
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++lt) {
                const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
                mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);        // epilogue has drained this accumulator (first two pass)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(full_bar(s), ph);                      // TMA landed this stage
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = base + s * STAGE_BYTES, b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (>> 4) start-address field
                        umma_bf16(tmem_d, make_desc(a_addr + k * UMMA_K * 2), make_desc(b_addr + k * UMMA_K * 2), C::IDESC, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));                       // stage reusable once these MMAs have read it
                }
                umma_commit(tmem_full_bar(acc));                     // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        // TMEM row -> registers -> bias / activation -> bf16 -> this warp's 32-row staging tile (16-byte chunks XOR-swizzled by
        // the row so both the row-per-lane writes and the row-major reads are conflict-free) -> global stores in which the
        // warp covers whole output rows.
        const int q = warp & 3;
        constexpr int ROW_BYTES = BN * 2, CHUNKS = BN / 8;   // 16-byte chunks per staged row
        uint8_t *stage = smem_gen + (out_stage - base) + q * (32 * ROW_BYTES);
        float *s_bias = reinterpret_cast<float *>(smem_gen + (bias_stage - base)) + q * BN;
        const bool vec_ok = (N % 8) == 0;
        uint32_t lt = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++lt) {
            const int m0 = (int)(tile / n_blocks) * BM, n0 = (int)(tile % n_blocks) * BN;
            const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
            // this warp's copy of the tile's bias slice (overlaps the wait for the accumulator)
#pragma unroll
            for (int j = lane; j < BN; j += 32) s_bias[j] = n0 + j < N ? __ldg(bias + n0 + j) : 0.f;
            __syncwarp();
            mbar_wait(tmem_full_bar(acc), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(s_bias + c + j), b1 = *reinterpret_cast<const float4 *>(s_bias + c + j + 4);
                    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float a = activate<ACT>(__uint_as_float(r[j + 2 * t]) + bv[2 * t], slope);
                        const float b = activate<ACT>(__uint_as_float(r[j + 2 * t + 1]) + bv[2 * t + 1], slope);
                        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                        pk[t] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    const int chunk = (c + j) >> 3;
                    *reinterpret_cast<uint4 *>(stage + lane * ROW_BYTES + ((chunk ^ (lane & (CHUNKS - 1))) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            // every TMEM read of this accumulator has completed: hand it back before writing the tile out
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
            constexpr int ROWS_PER_IT = 32 / CHUNKS;   // rows covered by one warp-wide 16-byte store (2 for BN 128, 1 for 256)
            const int chunk = lane % CHUNKS, col = n0 + chunk * 8, rsub = lane / CHUNKS;
            const int rows_here = min(32, M - (m0 + q * 32));   // may be <= 0
            if (vec_ok && col + 8 <= N) {
#pragma unroll 4
                for (int it = 0; it < 32 / ROWS_PER_IT; ++it) {
                    const int rl = ROWS_PER_IT * it + rsub;
                    if (rl < rows_here) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(stage + rl * ROW_BYTES + ((chunk ^ (rl & (CHUNKS - 1))) << 4));
                        *reinterpret_cast<uint4 *>(y + (size_t)(m0 + q * 32 + rl) * N + col) = v;
                    }
                }
            } else if (col < N) {
                for (int it = 0; it < 32 / ROWS_PER_IT; ++it) {
                    const int rl = ROWS_PER_IT * it + rsub;
                    if (rl >= rows_here) break;
                    const uint4 v = *reinterpret_cast<const uint4 *>(stage + rl * ROW_BYTES + ((chunk ^ (rl & (CHUNKS - 1))) << 4));
                    const __nv_bfloat16 *e = reinterpret_cast<const __nv_bfloat16 *>(&v);
                    __nv_bfloat16 *dst = y + (size_t)(m0 + q * 32 + rl) * N + col;
                    for (int t = 0; t < 8 && col + t < N; ++t) dst[t] = e[t];
                }
            }
            __syncwarp();   // the staging and bias tiles are rewritten by the next tile
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}




// ---------------------------------------------------------------------------------------------------------------------------
// CTA-pair variant of the dense layer for large shapes (tcgen05 cta_group::2, cluster of two CTAs on one TPC): the pair owns a
// 256 x 256 output tile; CTA r loads its own 128 rows of x and HALF of the weight rows (output columns [128 r, +128)) per K
// step (16 + 16 KB per stage, ring of six), the leader's elected thread issues tcgen05.mma.cta_group::2 M256 N256 K16 (A from
// each CTA's shared memory, the B halves from both), each CTA's accumulator is 128 lanes x 256 columns so two sets fit in TMEM
// and the epilogue (bias + activation, bf16, 2 KB blocks in TMA's 64-byte swizzle, one bulk tensor store per block; rows / columns
// beyond M / N are clipped by the tensor map) runs under the next tile's MMAs.  Barrier plumbing as in conv3x3_gn_pair_kernel.
// ---------------------------------------------------------------------------------------------------------------------------
namespace lpair {
constexpr int BN2 = 256, HALF_BYTES = 128 * BK * 2;     // 16 KB: 128 rows x 64 K of x, or of w
constexpr int STAGE_BYTES2 = 2 * HALF_BYTES, STAGES2 = 6;
constexpr int EPI_WARPS2 = 16, THREADS2 = 64 + 32 * EPI_WARPS2;
constexpr int OUT_BUF2 = 32 * 32 * 2, OUT_BYTES2 = EPI_WARPS2 * OUT_BUF2;
constexpr size_t SMEM_BYTES2 = 1024 + (size_t)STAGES2 * STAGE_BYTES2 + OUT_BYTES2 + 512;
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}   // namespace lpair

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(lpair::THREADS2, 1)
linear_bf16_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                        const __grid_constant__ CUtensorMap map_y, const float *__restrict__ bias, int M, int N, int K, float slope) {
    using namespace lpair;
    using namespace pairops;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_base = base + STAGES2 * STAGE_BYTES2;
    const uint32_t bars = out_base + OUT_BYTES2;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES2 + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (2 * STAGES2 + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (2 * STAGES2 + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES2 + 4);
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_blocks = (N + BN2 - 1) / BN2, m_blocks = (M + 255) / 256;
    const int tiles = n_blocks * m_blocks;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), 2 * EPI_WARPS2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer (both CTAs) =====
            uint32_t it = 0;
            for (int tile = pair_id; tile < tiles; tile += n_pairs) {
                const int m0 = (tile / n_blocks) * 256 + (int)rank * 128, n0 = (tile % n_blocks) * BN2 + (int)rank * 128;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STAGES2;
                    mbar_wait(empty_bar(s), ((it / STAGES2) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES2);
                    const uint32_t fb = map_to_cta(full_bar(s), 0);
                    tma2_load_2d(base + s * STAGE_BYTES2, &map_x, fb, kb * BK, m0);
                    tma2_load_2d(base + s * STAGE_BYTES2 + HALF_BYTES, &map_w, fb, kb * BK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader && lane == 0) {   // ===== MMA issuer (leader) =====
            uint32_t it = 0, lt = 0;
            for (int tile = pair_id; tile < tiles; tile += n_pairs, ++lt) {
                const uint32_t acc = lt & 1u;
                mbar_wait(tmem_empty(acc), ((lt >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * BN2;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STAGES2;
                    mbar_wait(full_bar(s), (it / STAGES2) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = base + s * STAGE_BYTES2, b_addr = a_addr + HALF_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma2_bf16(tmem_d, make_desc(a_addr + k * UMMA_K * 2), make_desc(b_addr + k * UMMA_K * 2), IDESC2, (kb | k) ? 1u : 0u);
                    umma2_commit(empty_bar(s));
                }
                umma2_commit(tmem_full(acc));
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue (both CTAs): own 128 rows; TMEM lane quarter = warp % 4, 64-column block = (warp - 2) / 4 =====
        const int q = warp & 3, part = (warp - 2) >> 2;
        const uint32_t buf = out_base + (uint32_t)(warp - 2) * OUT_BUF2;
        uint32_t lt = 0;
        for (int tile = pair_id; tile < tiles; tile += n_pairs, ++lt) {
            const int m0 = (tile / n_blocks) * 256 + (int)rank * 128, n0 = (tile % n_blocks) * BN2;
            const uint32_t acc = lt & 1u;
            mbar_wait(tmem_full(acc), (lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = part * 64; c < part * 64 + 64; c += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2 + c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                __syncwarp();
                const int col0 = n0 + c;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float bv[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) bv[t] = col0 + j + t < N ? __ldg(bias + col0 + j + t) : 0.f;
                    uint32_t pk[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float a = activate<ACT>(__uint_as_float(r[j + 2 * t]) + bv[2 * t], slope);
                        const float b = activate<ACT>(__uint_as_float(r[j + 2 * t + 1]) + bv[2 * t + 1], slope);
                        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                        pk[t] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    const uint32_t dst = buf + (uint32_t)lane * 64u + ((uint32_t)((j >> 3) ^ ((lane >> 1) & 3)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && col0 < N && m0 + q * 32 < M) {   // the tensor map clips rows >= M / columns >= N of the block
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&map_y), "r"(buf), "r"(col0), "r"(m0 + q * 32) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(tmem_empty(acc), 0));
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// y as a row-major [M][N] bf16 matrix; box = 32 columns x 32 rows in the 64-byte swizzle the pair kernel's epilogue writes
static bool make_out_map2(CUtensorMap *map, const void *ptr, int M, int N) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Stem of the stand-in backbone as an implicit GEMM on the tensor cores: the 7x7/2 convolution runs as a 4x4/1 convolution
// over the packed 2x2 space-to-depth image P (N, Hp, Wp, 16) (stem_s2d_pack), i.e. out[(n,y,x), o] = sum_{dy<4} sum_{j<64}
// P[n, y+dy, x .. x+3, :][j] * W[o, dy*64 + j]: four K = 64 blocks, each a row of 64 CONTIGUOUS elements of P.  A tensor map with
// overlapping rows (row pitch 16 elements = 32 bytes, row length 64) makes TMA do the im2col: the block of output row y and tap dy
// is the box {64, 128} at flat pixel (n*Hp + y+dy)*Wp -- and it is the SAME block output row y+1 needs for tap dy-1, so a CTA
// walking down an image loads ONE new 16 KB block per 128-pixel output row and keeps the last four in a ring (cuDNN's kernel
// for this layer runs at 0.2 PFLOP/s: 2.8 ms per 1024 RoIs).  The weights (64 x 256 bf16 = 32 KB) stay in shared memory for the
// whole kernel.  Roles as in linear_bf16_kernel: TMA producer warp, MMA warp (tcgen05.mma M128 N64 K16, fp32 accumulators double
// buffered in TMEM), four epilogue warps (bias + ReLU, bf16, row-contiguous stores).
// ---------------------------------------------------------------------------------------------------------------------------
namespace stem {
constexpr int BN = 64, RING = 8, ROWS_PER_ITEM = 32, TAPS = 4;
constexpr int BLK_BYTES = BM * BK * 2;              // one im2col block: 128 pixels x 64 elements
constexpr int W_BYTES = TAPS * BN * BK * 2;         // resident weights: 4 K-blocks of 64 rows x 128 bytes
constexpr int OUT_BYTES = BM * BN * 2;              // one activated output row: 128 pixels x 64 channels bf16
constexpr int OUT_ROWS = 3;                         // POOL: ring of the last three activated rows (3x3/2 max-pool window)
constexpr int TMEM_COLS = 2 * BN;
constexpr size_t SMEM_BYTES = 1024 + (size_t)RING * BLK_BYTES + W_BYTES + OUT_ROWS * OUT_BYTES + 4 * BN * 4 + 256;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}   // namespace stem

template <bool POOL>
__global__ void __launch_bounds__(THREADS, 1)
stem_s2d_gemm_kernel(const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_w, const float *__restrict__ bias,
                     __nv_bfloat16 *__restrict__ y, int n_img, int Hp, int Wp, int Ho) {
    using namespace stem;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = base + RING * BLK_BYTES;
    const uint32_t out_stage = w_base + W_BYTES;
    const uint32_t bias_stage = out_stage + OUT_ROWS * OUT_BYTES;
    const uint32_t bars = bias_stage + 4 * BN * 4;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (RING + s); };
    auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * RING + a); };
    auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * RING + 2 + a); };
    const uint32_t w_bar = bars + 8u * (2 * RING + 4);
    // work item = (image, chunk of ROWS_PER_ITEM output rows); with POOL the chunk also computes the row above it (the pooling
    // window of its first pooled row reaches one row up)
    auto item_rows = [&](int item, int &n, int &ys, int &rows) {
        const int chunks_ = (Ho + ROWS_PER_ITEM - 1) / ROWS_PER_ITEM;
        n = item / chunks_;
        const int y0 = (item - n * chunks_) * ROWS_PER_ITEM;
        ys = POOL && y0 > 0 ? y0 - 1 : y0;
        rows = min(y0 + ROWS_PER_ITEM, Ho) - ys;
    };
    const uint32_t tmem_slot = bars + 8u * (2 * RING + 5);
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = (Ho + ROWS_PER_ITEM - 1) / ROWS_PER_ITEM;        // work item = (image, chunk of ROWS_PER_ITEM output rows)
    const int items = n_img * chunks;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < RING; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 4);
        }
        mbar_init(w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        // ===== TMA producer: the weights once, then one im2col block per input row of every work item =====
        if (lane == 0) {
            mbar_expect_tx(w_bar, W_BYTES);
            for (int kb = 0; kb < TAPS; ++kb) tma_load_2d(w_base + kb * (BN * BK * 2), &map_w, w_bar, kb * BK, 0);
            uint32_t g = 0;   // blocks issued so far
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int n, y0, rows;
                item_rows(item, n, y0, rows);
                for (int r = 0; r < rows + TAPS - 1; ++r, ++g) {
                    const int s = g % RING;
                    const uint32_t ph = (g / RING) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), BLK_BYTES);
                    tma_load_2d(base + s * BLK_BYTES, &map_v, full_bar(s), 0, (n * Hp + y0 + r) * Wp);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            mbar_wait(w_bar, 0u);
            uint32_t g0 = 0, waited = 0, lt = 0;   // first block of the item / blocks whose arrival has been observed / tiles done
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int n, y0, rows;
                item_rows(item, n, y0, rows);
                (void)n; (void)y0;
                for (int t = 0; t < rows; ++t, ++lt) {
                    const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
                    mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);
                    while (waited <= g0 + t + TAPS - 1) {   // blocks arrive in issue order; each is waited for exactly once
                        mbar_wait(full_bar(waited % RING), (waited / RING) & 1u);
                        ++waited;
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + acc * BN;
#pragma unroll
                    for (int dy = 0; dy < TAPS; ++dy) {
                        const uint32_t a_addr = base + ((g0 + t + dy) % RING) * BLK_BYTES, b_addr = w_base + dy * (BN * BK * 2);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(tmem_d, make_desc(a_addr + k * UMMA_K * 2), make_desc(b_addr + k * UMMA_K * 2), IDESC, (dy | k) ? 1u : 0u);
                    }
                    umma_commit(empty_bar((g0 + t) % RING));          // input row y0+t is not needed by the rows below
                    if (t == rows - 1)
                        for (int e = 1; e < TAPS; ++e) umma_commit(empty_bar((g0 + t + e) % RING));
                    umma_commit(tmem_full_bar(acc));
                }
                g0 += rows + TAPS - 1;
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..5 =====
        // TMEM -> bias + ReLU -> bf16 -> the activated row in shared memory (pixel-major, 16-byte chunks XOR-swizzled by the pixel).
        // !POOL: the warp stores its 32 pixels.  POOL: the row goes into a ring of three; after every odd row y = 2 py + 1 the
        // four warps together write pooled row py = max over rows 2py-1 .. 2py+1 x columns 2px-1 .. 2px+1 (MaxPool2d(3, 2, 1);
        // ReLU commutes with max), so the pre-pool tensor never leaves the SM.
        const int q = warp & 3, et = threadIdx.x - 64;   // 0..127 among the epilogue threads
        constexpr int ROW_BYTES = BN * 2, CHUNKS = BN / 8;
        uint8_t *ring = smem_gen + (out_stage - base);
        float *s_bias = reinterpret_cast<float *>(smem_gen + (bias_stage - base)) + q * BN;
#pragma unroll
        for (int j = lane; j < BN; j += 32) s_bias[j] = __ldg(bias + j);
        __syncwarp();
        const int Hq = (Ho - 1) / 2 + 1, Wq = (BM - 1) / 2 + 1;   // pooled size
        uint32_t lt = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int n, ys, rows;
            item_rows(item, n, ys, rows);
            const int chunks_ = (Ho + ROWS_PER_ITEM - 1) / ROWS_PER_ITEM, y0 = (item - n * chunks_) * ROWS_PER_ITEM;
            for (int t = 0; t < rows; ++t, ++lt) {
                const int yrow = ys + t;
                const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
                uint8_t *stage = ring + (POOL ? (yrow % OUT_ROWS) * OUT_BYTES : 0) + q * (32 * ROW_BYTES);
                mbar_wait(tmem_full_bar(acc), acc_ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint32_t pk[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float a = fmaxf(__uint_as_float(r[j + 2 * u]) + s_bias[c + j + 2 * u], 0.f);
                            const float b = fmaxf(__uint_as_float(r[j + 2 * u + 1]) + s_bias[c + j + 2 * u + 1], 0.f);
                            const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                            pk[u] = *reinterpret_cast<const uint32_t *>(&h);
                        }
                        const int chunk = (c + j) >> 3;
                        *reinterpret_cast<uint4 *>(stage + lane * ROW_BYTES + ((chunk ^ (lane & (CHUNKS - 1))) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
                if (!POOL) {
                    // output row (n, yrow): 128 pixels x 64 channels contiguous; this warp writes pixels 32q .. 32q+31
                    __nv_bfloat16 *dst = y + (((size_t)n * Ho + yrow) * BM + q * 32) * BN;
                    constexpr int ROWS_PER_IT = 32 / CHUNKS;   // 4 pixels per warp-wide 16-byte store
                    const int chunk = lane % CHUNKS, rsub = lane / CHUNKS;
#pragma unroll
                    for (int it = 0; it < 32 / ROWS_PER_IT; ++it) {
                        const int rl = ROWS_PER_IT * it + rsub;
                        const uint4 v = *reinterpret_cast<const uint4 *>(stage + rl * ROW_BYTES + ((chunk ^ (rl & (CHUNKS - 1))) << 4));
                        *reinterpret_cast<uint4 *>(dst + (size_t)rl * BN + chunk * 8) = v;
                    }
                    __syncwarp();
                } else {
                    asm volatile("bar.sync 1, 128;" ::: "memory");   // row yrow staged by all four warps
                    const bool last = yrow == Ho - 1;
                    if (((yrow & 1) || last) && yrow >= y0) {
                        // pooled row py: odd yrow = 2py+1 closes its window; an even last row closes py = yrow/2 on its own
                        const int py = yrow >> 1;
                        if ((yrow & 1) ? (2 * py >= y0) : true) {
                            // branch-free 3x3 window: taps outside the map are replaced by the nearest tap inside (max is idempotent),
                            // so the nine 16-byte loads are unconditional and in flight together
                            const uint8_t *rw[3] = {ring + (max(2 * py - 1, 0) % OUT_ROWS) * OUT_BYTES, ring + ((2 * py) % OUT_ROWS) * OUT_BYTES,
                                                    ring + (min(2 * py + 1, Ho - 1) % OUT_ROWS) * OUT_BYTES};
                            __nv_bfloat16 *dst = y + ((size_t)n * Hq + py) * Wq * BN;
#pragma unroll 2
                            for (int id = et; id < Wq * CHUNKS; id += 128) {
                                const int px = id / CHUNKS, chunk = id - px * CHUNKS;
                                const int cc[3] = {max(2 * px - 1, 0), 2 * px, min(2 * px + 1, BM - 1)};
                                uint4 v[9];
#pragma unroll
                                for (int a = 0; a < 3; ++a)
#pragma unroll
                                    for (int b = 0; b < 3; ++b)
                                        v[a * 3 + b] = *reinterpret_cast<const uint4 *>(rw[a] + cc[b] * ROW_BYTES + ((chunk ^ (cc[b] & (CHUNKS - 1))) << 4));
                                __align__(16) __nv_bfloat162 m[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    m[u] = reinterpret_cast<const __nv_bfloat162 *>(&v[0])[u];
#pragma unroll
                                    for (int k = 1; k < 9; ++k) m[u] = __hmax2(m[u], reinterpret_cast<const __nv_bfloat162 *>(&v[k])[u]);
                                }
                                *reinterpret_cast<uint4 *>(dst + (size_t)px * BN + chunk * 8) = *reinterpret_cast<const uint4 *>(m);
                            }
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");   // the oldest ring slot is overwritten by the next row
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)stem::TMEM_COLS) : "memory");
    }
}

// overlapping-row view of the packed image: row r = 64 elements starting at flat pixel r (pitch 16 elements = 32 bytes)
static bool make_im2col_map(CUtensorMap *map, const void *ptr, long long n_pixels) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)BK, (cuuint64_t)(n_pixels - 3)};
    const cuuint64_t strides[1] = {(cuuint64_t)16 * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace gp

constexpr int kMaxDevices = 64;

template <int BN, int ACT>
static int launch_linear(const void *x, const void *w, const float *bias, void *y, int M, int N, int K, float slope, cudaStream_t st) {
    using namespace gp::tc;
    using C = Cfg<BN>;
    // the opt-in shared-memory size is a per-device function attribute and the grid is one CTA per SM of THAT device
    static int sms_of[kMaxDevices] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) return GP_ERR_UNSUPPORTED;
    if (!sms_of[dev]) {
        cudaError_t e = cudaFuncSetAttribute(linear_bf16_kernel<BN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
    }
    const int sms = sms_of[dev];
    CUtensorMap mx, mw;
    if (!make_map(&mx, x, M, K, BM) || !make_map(&mw, w, N, K, BN)) return GP_ERR_UNSUPPORTED;
    const long long tiles = (long long)((N + BN - 1) / BN) * ((M + BM - 1) / BM);
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);   // persistent: one CTA per SM
    linear_bf16_kernel<BN, ACT><<<grid, THREADS, C::SMEM_BYTES, st>>>(mx, mw, bias, (__nv_bfloat16 *)y, M, N, K, slope);
    gp::g_launches += 1;
    return (int)cudaGetLastError();
}

template <int ACT>
static int launch_linear_pair(const void *x, const void *w, const float *bias, void *y, int M, int N, int K, float slope, cudaStream_t st) {
    using namespace gp::tc;
    using namespace gp::tc::lpair;
    static int sms_of[kMaxDevices] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) return GP_ERR_UNSUPPORTED;
    if (!sms_of[dev]) {
        cudaError_t e = cudaFuncSetAttribute(linear_bf16_pair_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES2);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
    }
    CUtensorMap mx, mw, my;
    if (!make_map(&mx, x, M, K, 128) || !make_map(&mw, w, N, K, 128) || !make_out_map2(&my, y, M, N)) return GP_ERR_UNSUPPORTED;
    const long long tiles = (long long)((N + 255) / 256) * ((M + 255) / 256);
    const long long pairs = tiles < sms_of[dev] / 2 ? tiles : sms_of[dev] / 2;
    linear_bf16_pair_kernel<ACT><<<(unsigned)(2 * pairs), THREADS2, SMEM_BYTES2, st>>>(mx, mw, my, bias, M, N, K, slope);
    gp::g_launches += 1;
    return (int)cudaGetLastError();
}

// kernel choice for large shapes: -1 = automatic (CTA pair when the 256 x 256 tiles fill the machine), 0 = never, 1 = whenever legal
static int g_linear_pair = -1;

template <int BN>
static int launch_linear_act(const void *x, const void *w, const float *bias, void *y, int M, int N, int K, int act, float slope,
                             cudaStream_t st) {
    using namespace gp::tc;
    if (act == ACT_LRELU) return launch_linear<BN, ACT_LRELU>(x, w, bias, y, M, N, K, slope, st);
    if (act == ACT_RELU) return launch_linear<BN, ACT_RELU>(x, w, bias, y, M, N, K, slope, st);
    return launch_linear<BN, ACT_NONE>(x, w, bias, y, M, N, K, slope, st);
}

extern "C" int gp_linear_set_pair(int mode) {   // -1 / 2 automatic, 0 never, 1 whenever legal; returns the previous mode
    const int old = g_linear_pair < 0 ? 2 : g_linear_pair;
    g_linear_pair = mode < 0 ? 2 : mode;
    return old;
}

extern "C" int gp_linear_bf16(const void *x, const void *w, const float *bias, void *y, int M, int N, int K, int act, float slope,
                              void *stream) {
    if (!x || !w || !bias || !y) return GP_ERR_NULL;
    if (M < 0 || N <= 0 || K <= 0 || K % 8) return GP_ERR_SHAPE;   // 16-byte row pitch for the tensor maps
    if (act < 0 || act > 2 || (act == 1 && !(slope >= 0.f && slope <= 1.f))) return GP_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15u) return GP_ERR_ALIGN;
    if (M == 0) return GP_OK;
    // CTA pairs (cta_group::2, 256 x 256 tiles) for the large layers: N a multiple of 8 (16-byte output rows for the TMA store)
    {
        if (g_linear_pair < 0) { const char *e = getenv("GP_LINEAR_PAIR"); g_linear_pair = e ? atoi(e) : 2; }
        const long long t256 = (long long)((N + 255) / 256) * ((M + 255) / 256);
        const bool legal = N % 8 == 0 && M >= 128 && N >= 128;
        if (legal && (g_linear_pair == 1 || (g_linear_pair == 2 && t256 >= 64 && K >= 512))) {
            cudaStream_t st = (cudaStream_t)stream;
            if (act == gp::tc::ACT_LRELU) return launch_linear_pair<gp::tc::ACT_LRELU>(x, w, bias, y, M, N, K, slope, st);
            if (act == gp::tc::ACT_RELU) return launch_linear_pair<gp::tc::ACT_RELU>(x, w, bias, y, M, N, K, slope, st);
            return launch_linear_pair<gp::tc::ACT_NONE>(x, w, bias, y, M, N, K, slope, st);
        }
    }
    // 128 x 256 tiles when there are enough of them to fill the machine: x is read once per 256 output columns, and a
    // K16 step reads 12 KB of operands per 128 tensor-pipe cycles instead of 8 KB per 64 (shared-memory bound at N = 128)
    const long long tiles256 = (long long)((N + 255) / 256) * ((M + 127) / 128);
    if (N > 128 && tiles256 >= 148) return launch_linear_act<256>(x, w, bias, y, M, N, K, act, slope, (cudaStream_t)stream);
    return launch_linear_act<128>(x, w, bias, y, M, N, K, act, slope, (cudaStream_t)stream);
}

extern "C" int gp_stem_s2d_gemm(const void *packed, const void *w, const float *bias, void *y, int N, int Hp, int Wp, int pool, void *stream) {
    using namespace gp::tc;
    if (!packed || !w || !bias || !y) return GP_ERR_NULL;
    const int Ho = Hp - 3, Wo = Wp - 3;
    if (N <= 0 || Ho <= 0 || Wo != BM) return GP_ERR_UNSUPPORTED;   // one 128-pixel output row per tile (256 x 256 crops)
    if ((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15u) return GP_ERR_ALIGN;
    if ((long long)N * Hp * Wp >= (1ll << 31)) return GP_ERR_SHAPE;
    static int sms_of[kMaxDevices] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) return GP_ERR_UNSUPPORTED;
    if (!sms_of[dev]) {
        cudaError_t e = cudaFuncSetAttribute(stem_s2d_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_s2d_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
    }
    const int sms = sms_of[dev];
    CUtensorMap mv, mw;
    if (!make_im2col_map(&mv, packed, (long long)N * Hp * Wp) || !make_map(&mw, w, stem::BN, stem::TAPS * BK, stem::BN)) return GP_ERR_UNSUPPORTED;
    const int items = N * ((Ho + stem::ROWS_PER_ITEM - 1) / stem::ROWS_PER_ITEM);
    const unsigned grid = (unsigned)(items < sms ? items : sms);
    if (pool) stem_s2d_gemm_kernel<true><<<grid, THREADS, stem::SMEM_BYTES, (cudaStream_t)stream>>>(mv, mw, bias, (__nv_bfloat16 *)y, N, Hp, Wp, Ho);
    else stem_s2d_gemm_kernel<false><<<grid, THREADS, stem::SMEM_BYTES, (cudaStream_t)stream>>>(mv, mw, bias, (__nv_bfloat16 *)y, N, Hp, Wp, Ho);
    gp::g_launches += 1;
    return (int)cudaGetLastError();
}

