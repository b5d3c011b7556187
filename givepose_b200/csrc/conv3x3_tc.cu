// givepose_b200 -- the coordinate-map decoder's 3x3 convolutions as a hand-written tcgen05 implicit GEMM (sm_100a).
//
// Replaces the cuDNN convolutions of TopDownXyzHead's ConvModules (network/xyz_head.py:195-366: six Conv2d(256, 256, 3, padding=1,
// bias=False) per head at 16x16, 32x32 and 64x64, each followed by GroupNorm(32) + GELU, network/torch_utils/layers/
// conv_module.py:57-234) -- 96 % of the decoder's FLOPs -- and produces the GroupNorm statistics in the same pass:
//
//     y[(n,h,w), o] = sum_{ky,kx,c} x[n, h+ky-1, w+kx-1, c] * wgt[o, ky, kx, c]       bf16 operands, fp32 accumulation in TMEM
//     partial[n][slab][g][0..1] = (sum, sum of squares) of the fp32 accumulators of group g over the slab's 64 pixels
//
// Implicit GEMM: M = N*H*W output pixels, N = 256 output channels, K = 9 taps x Cin.  No im2col buffer exists anywhere: for a tile
// of 256 output pixels (256/W image rows) and one (kx, 64-channel block) pair, ONE 4-D TMA box {64 channels, W, 256/W + 2 rows,
// 1 image} at coordinates {64 kc, kx-1, h0-1, n} of the channel-last activation brings the input slab in; coordinates that fall
// outside the image (-1 or H / W) are zero-filled by TMA, which is exactly the zero padding.  The slab lands in shared memory as
// rows of 128 bytes (one pixel x 64 channels) with the 128-byte swizzle the tensor core's K-major descriptor expects, and because
// one image row is W x 128 bytes = a multiple of the 1024-byte swizzle atom (W >= 8), the A operand of vertical tap ky for the
// tile's accumulator `sub` is simply the slab at byte offset (sub * 128/W + ky) * W * 128: the three vertical taps re-use the same
// shared-memory data, which halves the activation traffic through L2 (the first version of this kernel re-loaded every tap and
// ran at the L2 -> SM throughput cap: 64 KB of operands per 1024 tensor-pipe cycles; this one needs 48 KB per 1024).
//
// CTA tile = 256 pixels x 256 channels: two M128 x N256 accumulators = all 512 TMEM columns, so every 32 KB weight block
// (256 rows x 64 K) that comes through shared memory feeds 2 x 4 MMAs.  Persistent CTAs (one per SM, 576 threads, warp-specialised):
//   warp 0     TMA producer: per (kx, kc) one activation slab into a ring of two, per tap one 2-D weight box into a ring of four
//              32 KB stages; completion on mbarriers (expect_tx); runs ahead across tile boundaries
//   warp 1     TMEM allocation (512 columns); one elected thread issues tcgen05.mma.cta_group::1.kind::f16 M128 N256 K16,
//              tcgen05.commit releases the weight stage / the slab / publishes the accumulators
//   warps 2-17 epilogue (four warps per TMEM lane quarter, 64 columns each): tcgen05.ld 32 lanes x 32 columns, GroupNorm partial
//              sums of the fp32 values (transposing butterfly over the 32 rows of the warp: 9 shuffles per 4 groups), bf16 pack
//              into a swizzled 2 KB staging block, one TMA bulk tensor store per block; the partials go out in the [n][slab][g][2] layout gn_finalize_kernel sums in a fixed order
//              (no atomics: bit-reproducible)
#include <stdlib.h>

#include "tc_common.cuh"
#include "gelu_fast.cuh"

#ifndef GP_CONV_PAIR_DEFAULT
#define GP_CONV_PAIR_DEFAULT 1
#endif

namespace gp {
namespace tc {
namespace conv {

constexpr int BN = 256;                          // output channels = the whole Cout of the decoder
constexpr int SUB = 2;                           // M128 sub-tiles (accumulators) per CTA tile
constexpr int TILE_PIX = SUB * BM;               // 256 output pixels per tile
constexpr int B_BYTES = BN * BK * 2;             // 32 KB weight block: one tap x 64 input channels x 256 output channels
constexpr int B_STAGES = 3;
constexpr int MAX_W = 64;                        // widest image row: the slab holds 256/W + 2 rows of W pixels
constexpr int SLAB_BYTES = (TILE_PIX + 2 * MAX_W) * BK * 2;   // 48 KB
constexpr int SLABS = 2;
constexpr int EPI_WARPS = 16;                    // four per TMEM lane quarter, 64 accumulator columns each
constexpr int CONV_THREADS = 64 + 32 * EPI_WARPS; // producer warp, MMA warp, epilogue warps
constexpr int TMEM_COLS = SUB * BN;              // 512: the whole tensor memory of the SM
constexpr int OUT_BUF_BYTES = 32 * 32 * 2;       // epilogue staging: 32 rows x 32 channels bf16 per TMA store, one buffer per warp
constexpr int OUT_BYTES = EPI_WARPS * OUT_BUF_BYTES;   // 32 KB
constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)SLABS * SLAB_BYTES + (size_t)B_STAGES * B_BYTES + OUT_BYTES + 256 /*barriers*/;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr int GN_GROUPS = 32;                    // GroupNorm(32, 256): 8 channels per group

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// Sum eight per-lane values over the 32 lanes with a transposing butterfly (4 + 2 + 1 + 2 shuffles).  On return every lane
// holds the warp total of value index ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1).
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane) {
    const unsigned full = 0xffffffffu;
    float a[4], b[2], c;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = h16 ? v[4 + i] : v[i], send = h16 ? v[i] : v[4 + i];
        a[i] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = h8 ? a[2 + i] : a[i], send = h8 ? a[i] : a[2 + i];
        b[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    {
        const float keep = h4 ? b[1] : b[0], send = h4 ? b[0] : b[1];
        c = keep + __shfl_xor_sync(full, send, 4);
    }
    c += __shfl_xor_sync(full, c, 2);
    c += __shfl_xor_sync(full, c, 1);
    return c;
}

// tiles_per_img = H*W / 256; rows_per_sub = 128 / W image rows per accumulator; kc_blocks = Cin / 64; row_bytes = W * 128.
template <bool STATS>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_gn_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_y, float *__restrict__ partial /*[N][tiles_per_img*4][32][2]*/, int n_tiles,
                  int tiles_per_img, int rows_per_sub, int kc_blocks, int row_bytes) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
    const uint32_t b_base = base + SLABS * SLAB_BYTES;
    const uint32_t out_base = b_base + B_STAGES * B_BYTES;
    const uint32_t bars = out_base + OUT_BYTES;
    auto b_full = [&](int s) { return bars + 8u * s; };
    auto b_empty = [&](int s) { return bars + 8u * (B_STAGES + s); };
    auto a_full = [&](int s) { return bars + 8u * (2 * B_STAGES + s); };
    auto a_empty = [&](int s) { return bars + 8u * (2 * B_STAGES + SLABS + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * B_STAGES + 2 * SLABS), tmem_empty_bar = tmem_full_bar + 8u;
    const uint32_t tmem_slot = tmem_full_bar + 16u;
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = 3 * kc_blocks;                                   // (kx, kc) pairs per tile: one slab each
    const uint32_t slab_bytes = (uint32_t)(SUB * rows_per_sub + 2) * (uint32_t)row_bytes;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < B_STAGES; ++s) {
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), 1);
        }
        for (int s = 0; s < SLABS; ++s) {
            mbar_init(a_full(s), 1);
            mbar_init(a_empty(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, EPI_WARPS);   // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
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
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t ia = 0, ib = 0;   // slabs / weight blocks issued so far, across tiles
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int n = tile / tiles_per_img, h0 = (tile - n * tiles_per_img) * (SUB * rows_per_sub);
                for (int g = 0; g < groups; ++g, ++ia) {
                    const int kx = g / kc_blocks, kc = g - kx * kc_blocks;
                    const int sa = ia % SLABS;
                    mbar_wait(a_empty(sa), ((ia / SLABS) & 1u) ^ 1u);
                    mbar_expect_tx(a_full(sa), slab_bytes);
                    tma_load_4d(base + sa * SLAB_BYTES, &map_x, a_full(sa), kc * BK, kx - 1, h0 - 1, n);
                    for (int ky = 0; ky < 3; ++ky, ++ib) {
                        const int sb = ib % B_STAGES;
                        mbar_wait(b_empty(sb), ((ib / B_STAGES) & 1u) ^ 1u);
                        mbar_expect_tx(b_full(sb), B_BYTES);
                        tma_load_2d(b_base + sb * B_BYTES, &map_w, b_full(sb), ((ky * 3 + kx) * kc_blocks + kc) * BK, 0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t ia = 0, ib = 0, lt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                mbar_wait(tmem_empty_bar, (lt & 1u) ^ 1u);          // the epilogue has drained the previous tile (first passes)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int g = 0; g < groups; ++g, ++ia) {
                    const int sa = ia % SLABS;
                    mbar_wait(a_full(sa), (ia / SLABS) & 1u);
                    const uint32_t slab = base + sa * SLAB_BYTES;
                    for (int ky = 0; ky < 3; ++ky, ++ib) {
                        const int sb = ib % B_STAGES;
                        mbar_wait(b_full(sb), (ib / B_STAGES) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_addr = b_base + sb * B_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t bd = make_desc(b_addr + k * UMMA_K * 2);
#pragma unroll
                            for (int sub = 0; sub < SUB; ++sub)   // vertical tap ky of accumulator `sub`: the slab, (sub * rows + ky) image rows down
                                umma_bf16(tmem_base + sub * BN, make_desc(slab + (uint32_t)(sub * rows_per_sub + ky) * (uint32_t)row_bytes + k * UMMA_K * 2),
                                          bd, IDESC, (g | ky | k) ? 1u : 0u);
                        }
                        umma_commit(b_empty(sb));
                    }
                    umma_commit(a_empty(sa));
                }
                umma_commit(tmem_full_bar);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..17; TMEM lane quarter = warp % 4, column block = (warp - 2) / 4 =====
        // Output path: a lane holds 32 consecutive channels of ONE pixel, so direct stores would touch 32 different lines per
        // instruction (measured: the store wavefronts made the epilogue 13k cycles per tile).  Each 32 x 32 block goes through a
        // 2 KB shared-memory buffer in TMA's 64-byte swizzle (conflict-free 16-byte writes) and leaves as one bulk tensor store.
        const int q = warp & 3, part = (warp - 2) >> 2;
        constexpr int COLS = BN / (EPI_WARPS / 4);   // accumulator columns per warp
        const uint32_t buf = out_base + (uint32_t)(warp - 2) * OUT_BUF_BYTES;
        uint32_t lt = 0, nst = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int n = tile / tiles_per_img, t_in = tile - n * tiles_per_img;
            mbar_wait(tmem_full_bar, lt & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = part * COLS; c < (part + 1) * COLS; c += 32) {
                float st8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) st8[i] = 0.f;
#pragma unroll
                for (int sub = 0; sub < SUB; ++sub, ++nst) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * BN + c);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    // the bulk store of the previous block has finished reading the staging buffer
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    __syncwarp();
                    // this lane's output pixel: row 32q + lane of sub-tile `sub`; 32 consecutive channels c .. c+31
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint32_t pk[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float a = __uint_as_float(r[j + 2 * t]), b = __uint_as_float(r[j + 2 * t + 1]);
                            const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                            pk[t] = *reinterpret_cast<const uint32_t *>(&h);
                        }
                        const uint32_t dst = buf + (uint32_t)lane * 64u + ((uint32_t)((j >> 3) ^ ((lane >> 1) & 3)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                        if (STATS) {
                            float s = 0.f, ss = 0.f;
#pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                const float v = __uint_as_float(r[j + t]);
                                s += v;
                                ss = fmaf(v, v, ss);
                            }
                            st8[(j >> 3) * 2] += s;
                            st8[(j >> 3) * 2 + 1] += ss;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA store
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&map_y), "r"(buf), "r"(c), "r"(tile * TILE_PIX + sub * BM + q * 32) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (STATS) {
                    const float tot = warp_sum8(st8, lane);
                    if ((lane & 3) == 0) {
                        const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        const int g = (c >> 3) + (idx >> 1);
                        partial[(((size_t)n * (tiles_per_img * 4) + t_in * 4 + q) * GN_GROUPS + g) * 2 + (idx & 1)] = tot;
                    }
                }
            }
            // every TMEM read of this warp's part of the tile has completed: hand the accumulators back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory stays valid until the stores have read it
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2, cluster of two CTAs on one TPC).  The pair shares one 256-pixel x 256-channel tile:
// CTA r owns pixels [128 r, 128 r + 128) -- its own activation slab (128/W + 2 image rows) and HALF of every weight block (output
// channels [128 r, +128), 16 KB) -- and the leader's single thread issues tcgen05.mma.cta_group::2 M256 N256 K16, which reads A
// from each CTA's slab and the two B halves from both shared memories.  Per SM and K16 step that is 4 KB + 4 KB of operand reads
// instead of 4 KB + 8 KB, and 26 KB instead of 48 KB of TMA fills per K64 block, so the shared-memory port (128 B/clk) no longer
// limits the tensor pipe; and each CTA's accumulator is 128 lanes x 256 columns, so TWO accumulator sets fit in TMEM and the
// epilogue of tile i overlaps the MMAs of tile i+1.
//   full barriers live in the LEADER (its producer arms them for the bytes of both CTAs; the peer's TMA signals them through
//   .cta_group::2), empty / accumulator-full barriers are local and receive the leader's multicast tcgen05.commit, the
//   accumulator-empty barriers live in the leader and collect one arrival per epilogue warp of BOTH CTAs (remote arrive).
// ---------------------------------------------------------------------------------------------------------------------------
namespace pair {
constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;          // 16 KB: this CTA's half of a weight block
#ifndef GP_P_BSTAGES
#define GP_P_BSTAGES 6
#endif
#ifndef GP_P_SLABS
#define GP_P_SLABS 3
#endif
constexpr int P_B_STAGES = GP_P_BSTAGES;
constexpr int P_SLAB_BYTES = (BM + 2 * MAX_W) * BK * 2;     // 32 KB: 128/W + 2 rows of W pixels
constexpr int P_SLABS = GP_P_SLABS;
constexpr int ACCS = 2;                                   // accumulator sets (256 TMEM columns each)
constexpr size_t P_SMEM_BYTES = 1024 + (size_t)P_SLABS * P_SLAB_BYTES + (size_t)P_B_STAGES * B_HALF_BYTES + OUT_BYTES + 512;
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

using namespace gp::tc::pairops;
}   // namespace pair

// n_tiles 256-pixel tiles, one per CTA pair and step; rows_per_sub = 128 / W; partial: [N][tiles_per_img * 8][32][2]
//
// XFORM: the input x is the RAW output of the previous ConvModule's convolution, and this kernel applies that module's
// GroupNorm + GELU (conv_module.py order conv -> norm -> act) to the operand ON ITS WAY to the tensor core, so the apply pass
// over the previous activation (one read + one write of the whole tensor at the HBM rate) disappears: eight transform warps
// per CTA wait for a slab to land, rewrite it in place -- z = bf16(gelu(x * sc + sh)) with (sc, sh) folded from the producer
// layer's (mean, rstd), gamma, beta exactly as gn_apply_kernel does, so the operand bits are those the unfused pipeline would
// have read -- skip the positions TMA zero-filled (the convolution pads the ACTIVATED tensor with zeros), make the writes
// visible to the async proxy and arrive on the leader's `ready` barrier the MMA thread waits on.  A thread's 16-byte slots all
// hold the same channel octet (the slab's 128-byte swizzle depends on the pixel row mod 8 and a thread's rows are 32 apart), so
// its eight (sc, sh) pairs are loaded once per slab.  Each slab is one horizontal tap of one 64-channel block: an element is
// transformed three times per tile instead of once -- ALU work the MMA-bound kernel has issue slots for.
template <bool STATS, bool XFORM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1)
conv3x3_gn_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                       const __grid_constant__ CUtensorMap map_y, float *__restrict__ partial, int n_tiles, int tiles_per_img,
                       int rows_per_sub, int kc_blocks, int row_bytes, const float *__restrict__ in_stats /*[N][32][2]*/,
                       const float *__restrict__ in_gamma, const float *__restrict__ in_beta, int H, int lgW) {
    using namespace pair;
#ifndef GP_PAIR_EPI
#define GP_PAIR_EPI 16
#endif
    constexpr int NEPI = XFORM ? 8 : GP_PAIR_EPI;   // epilogue warps (XFORM: warps 2..9, transform warps 10..17)
    constexpr int NXF = 8;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = base + P_SLABS * P_SLAB_BYTES;
    const uint32_t out_base = b_base + P_B_STAGES * B_HALF_BYTES;
    const uint32_t bars = out_base + OUT_BYTES;
    auto b_full = [&](int s) { return bars + 8u * s; };
    auto b_empty = [&](int s) { return bars + 8u * (P_B_STAGES + s); };
    auto a_full = [&](int s) { return bars + 8u * (2 * P_B_STAGES + s); };
    auto a_empty = [&](int s) { return bars + 8u * (2 * P_B_STAGES + P_SLABS + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (2 * P_B_STAGES + 2 * P_SLABS + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (2 * P_B_STAGES + 2 * P_SLABS + ACCS + a); };
    auto a_ready = [&](int s) { return bars + 8u * (2 * P_B_STAGES + 2 * P_SLABS + 2 * ACCS + s); };   // XFORM: leader, 2 x NXF arrivals
    const uint32_t tmem_slot = bars + 8u * (2 * P_B_STAGES + 3 * P_SLABS + 2 * ACCS);
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int groups = 3 * kc_blocks;
    const uint32_t slab_bytes = (uint32_t)(rows_per_sub + 2) * (uint32_t)row_bytes;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P_B_STAGES; ++s) {
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), 1);
        }
        for (int s = 0; s < P_SLABS; ++s) {
            mbar_init(a_full(s), 1);
            mbar_init(a_empty(s), 1);
            mbar_init(a_ready(s), 2 * NXF);            // XFORM: every transform warp of both CTAs
        }
        for (int a = 0; a < ACCS; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), 2 * NEPI);        // every epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    }
    if (warp == 1) {   // one warp of EACH CTA takes part in the pair-wide allocation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers are initialised before any remote signal
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own slab + own half of the weight block, signalled on the LEADER's full barriers =====
        if (lane == 0) {
            uint32_t ia = 0, ib = 0;
            for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
                const int n = tile / tiles_per_img, h0 = (tile - n * tiles_per_img) * (SUB * rows_per_sub) + (int)rank * rows_per_sub;
                for (int g = 0; g < groups; ++g, ++ia) {
                    const int kx = g / kc_blocks, kc = g - kx * kc_blocks;
                    const int sa = ia % P_SLABS;
                    mbar_wait(a_empty(sa), ((ia / P_SLABS) & 1u) ^ 1u);
                    if (XFORM) {   // the slab lands on this CTA's own barrier: its transform warps take it from there
                        mbar_expect_tx(a_full(sa), slab_bytes);
                        tma_load_4d(base + sa * P_SLAB_BYTES, &map_x, a_full(sa), kc * BK, kx - 1, h0 - 1, n);
                    } else {
                        if (leader) mbar_expect_tx(a_full(sa), 2 * slab_bytes);
                        tma2_load_4d(base + sa * P_SLAB_BYTES, &map_x, map_to_cta(a_full(sa), 0), kc * BK, kx - 1, h0 - 1, n);
                    }
                    for (int ky = 0; ky < 3; ++ky, ++ib) {
                        const int sb = ib % P_B_STAGES;
                        mbar_wait(b_empty(sb), ((ib / P_B_STAGES) & 1u) ^ 1u);
                        if (leader) mbar_expect_tx(b_full(sb), 2 * B_HALF_BYTES);
                        tma2_load_2d(b_base + sb * B_HALF_BYTES, &map_w, map_to_cta(b_full(sb), 0), ((ky * 3 + kx) * kc_blocks + kc) * BK,
                                     (int)rank * (BN / 2));
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's elected thread, for the pair =====
        if (leader && lane == 0) {
            uint32_t ia = 0, ib = 0, lt = 0;
            for (int tile = pair_id; tile < n_tiles; tile += n_pairs, ++lt) {
                const uint32_t acc = lt & 1u;
                mbar_wait(tmem_empty(acc), ((lt >> 1) & 1u) ^ 1u);   // both CTAs have drained this accumulator set (first two pass)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int g = 0; g < groups; ++g, ++ia) {
                    const int sa = ia % P_SLABS;
                    mbar_wait(XFORM ? a_ready(sa) : a_full(sa), (ia / P_SLABS) & 1u);
                    const uint32_t slab = base + sa * P_SLAB_BYTES;
                    for (int ky = 0; ky < 3; ++ky, ++ib) {
                        const int sb = ib % P_B_STAGES;
                        mbar_wait(b_full(sb), (ib / P_B_STAGES) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_addr = b_base + sb * B_HALF_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma2_bf16(tmem_d, make_desc(slab + (uint32_t)ky * (uint32_t)row_bytes + k * UMMA_K * 2), make_desc(b_addr + k * UMMA_K * 2),
                                       IDESC2, (g | ky | k) ? 1u : 0u);
                        umma2_commit(b_empty(sb));
                    }
                    umma2_commit(a_empty(sa));
                }
                umma2_commit(tmem_full(acc));
            }
        }
        __syncwarp();
    } else if (XFORM && warp >= 2 + NEPI) {
        // ===== operand transform (both CTAs): GroupNorm + GELU of the producer layer, in place, slab by slab =====
        const int tx = threadIdx.x - 32 * (2 + NEPI);            // 0 .. 255
        const int j = tx & 7, jl = j ^ ((tx >> 3) & 7);          // physical 16-byte slot in the row / the channel octet it holds
        const int W = 1 << lgW, cpg = kc_blocks * BK / GN_GROUPS;  // channels per group of the producer's GroupNorm(32)
        const int npix = (rows_per_sub + 2) << lgW;
        uint8_t *slab0 = smem_gen;                                 // generic pointer to the slab ring
        uint32_t ia = 0;
        for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
            const int n = tile / tiles_per_img, h0 = (tile - n * tiles_per_img) * (SUB * rows_per_sub) + (int)rank * rows_per_sub;
            for (int g = 0; g < groups; ++g, ++ia) {
                const int kx = g / kc_blocks, kc = g - kx * kc_blocks;
                const int sa = ia % P_SLABS;
                // this thread's eight (scale, shift) pairs: channels c0 .. c0+7 of image n (one GroupNorm group when cpg % 8 == 0)
                const int c0 = kc * BK + jl * 8;
                const float mean = __ldg(in_stats + ((size_t)n * GN_GROUPS + c0 / cpg) * 2), rstd = __ldg(in_stats + ((size_t)n * GN_GROUPS + c0 / cpg) * 2 + 1);
                float sc[8], sh[8];
                {
                    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(in_gamma + c0)), g1 = __ldg(reinterpret_cast<const float4 *>(in_gamma + c0 + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(in_beta + c0)), b1 = __ldg(reinterpret_cast<const float4 *>(in_beta + c0 + 4));
                    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int t = 0; t < 8; ++t) {   // gn_fold: y = x * sc + sh
                        sc[t] = rstd * gm[t];
                        sh[t] = bt[t] - mean * sc[t];
                    }
                }
                mbar_wait(a_full(sa), (ia / P_SLABS) & 1u);                  // the slab has landed (TMA, async proxy)
                uint8_t *slab = slab0 + sa * P_SLAB_BYTES;
                for (int q = tx; q < npix * 8; q += 32 * NXF) {
                    const int r = q >> 3, row = r >> lgW, col = r & (W - 1);
                    const int ih = h0 - 1 + row, iw = col + kx - 1;
                    if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;      // zero padding of the ACTIVATED tensor: leave TMA's zeros
                    uint4 *cell = reinterpret_cast<uint4 *>(slab + (size_t)r * 128 + j * 16);
                    const uint4 v = *cell;
                    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
                    uint32_t o[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float x0 = __uint_as_float(wds[t] << 16), x1 = __uint_as_float(wds[t] & 0xffff0000u);
                        const float z0 = gelu_fast16(fmaf(x0, sc[2 * t], sh[2 * t])), z1 = gelu_fast16(fmaf(x1, sc[2 * t + 1], sh[2 * t + 1]));
                        const __nv_bfloat162 h = __floats2bfloat162_rn(z0, z1);
                        o[t] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    *cell = make_uint4(o[0], o[1], o[2], o[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(a_ready(sa), 0));
            }
        }
    } else if (warp < 2 + NEPI) {
        // ===== epilogue (both CTAs): own 128 pixels; TMEM lane quarter = warp % 4, column block = (warp - 2) / 4 =====
        const int q = warp & 3, part = (warp - 2) >> 2;
        constexpr int COLS = BN / (NEPI / 4);
        const uint32_t buf = out_base + (uint32_t)(warp - 2) * OUT_BUF_BYTES;
        uint32_t lt = 0;
        for (int tile = pair_id; tile < n_tiles; tile += n_pairs, ++lt) {
            const int n = tile / tiles_per_img, t_in = tile - n * tiles_per_img;
            const uint32_t acc = lt & 1u;
            mbar_wait(tmem_full(acc), (lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = part * COLS; c < (part + 1) * COLS; c += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c);
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
                float st8[8];
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float a = __uint_as_float(r[j + 2 * t]), b = __uint_as_float(r[j + 2 * t + 1]);
                        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                        pk[t] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    const uint32_t dst = buf + (uint32_t)lane * 64u + ((uint32_t)((j >> 3) ^ ((lane >> 1) & 3)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    float s = 0.f, ss = 0.f;
                    if (STATS) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const float v = __uint_as_float(r[j + t]);
                            s += v;
                            ss = fmaf(v, v, ss);
                        }
                    }
                    st8[(j >> 3) * 2] = s;
                    st8[(j >> 3) * 2 + 1] = ss;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&map_y), "r"(buf), "r"(c), "r"(tile * TILE_PIX + (int)rank * BM + q * 32) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (STATS) {
                    const float tot = warp_sum8(st8, lane);
                    if ((lane & 3) == 0) {
                        const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        const int g = (c >> 3) + (idx >> 1);
                        partial[(((size_t)n * (tiles_per_img * 8) + t_in * 8 + (int)rank * 4 + q) * GN_GROUPS + g) * 2 + (idx & 1)] = tot;
                    }
                }
            }
            // this warp's TMEM reads of the accumulator set are complete: tell the leader's MMA thread
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(tmem_empty(acc), 0));
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // nobody frees tensor memory / exits while the peer may still signal or read
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// channel-last activation (N, H, W, C) bf16 as a 4-D tensor {C, W, H, N}; box = {64 channels, W, rows, 1}: rows * W pixels x 128 bytes
static bool make_act_map(CUtensorMap *map, const void *ptr, int N, int H, int W, int C, int rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)W, (cuuint32_t)rows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// y as a row-major [pixels][256] bf16 matrix; box = 32 channels x 32 pixels in the 64-byte swizzle the epilogue writes
static bool make_out_map(CUtensorMap *map, const void *ptr, long long rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)BN, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)BN * 2};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace conv
}  // namespace tc
}  // namespace gp

// kernel variant: 0 = one CTA per tile (conv3x3_gn_kernel), 1 = CTA pair, tcgen05 cta_group::2 (conv3x3_gn_pair_kernel)
static int g_conv_pair = -1;
static int conv_pair_mode() {
    if (g_conv_pair < 0) {
        const char *e = getenv("GP_CONV_PAIR");
        g_conv_pair = e ? (atoi(e) ? 1 : 0) : GP_CONV_PAIR_DEFAULT;
    }
    return g_conv_pair;
}

extern "C" {

int gp_conv3x3_set_pair(int on) {
    const int old = conv_pair_mode();
    g_conv_pair = on ? 1 : 0;
    return old;
}

size_t gp_conv3x3_gn_slabs(int H, int W) {
    if (H <= 0 || W <= 0 || (H * W) % gp::tc::conv::TILE_PIX) return 0;
    return (size_t)(H * W / gp::tc::conv::TILE_PIX) * (conv_pair_mode() ? 8 : 4);
}

static int conv3x3_impl(const void *x, const void *w_packed, void *y, float *partial, int N, int H, int W, int Cin, int Cout, void *stream,
                        const float *in_stats, const float *in_gamma, const float *in_beta) {
    using namespace gp::tc;
    using namespace gp::tc::conv;
    if (!x || !w_packed || !y) return GP_ERR_NULL;
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0) return GP_ERR_SHAPE;
    const bool xform = in_stats != nullptr;
    if (xform) {   // fused GroupNorm(32) + GELU of the producer layer: CTA-pair kernel only, a 16-byte slot = channels of ONE group
        if (!in_gamma || !in_beta) return GP_ERR_NULL;
        if (!conv_pair_mode() || (Cin / 32) % 8 || Cin % 32) return GP_ERR_UNSUPPORTED;
        if ((reinterpret_cast<uintptr_t>(in_gamma) | reinterpret_cast<uintptr_t>(in_beta)) & 15u) return GP_ERR_ALIGN;
    }
    // 128 output pixels = whole image rows (W divides 128), a 256-pixel tile never straddles two images
    // and a shift by one image row inside the slab stays aligned to the 1024-byte swizzle atom (W >= 8)
    if (Cout != BN || Cin % BK || W > MAX_W || W < 8 || BM % W || (H * W) % TILE_PIX || H % (SUB * (BM / W))) return GP_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(y)) & 15u) return GP_ERR_ALIGN;
    const long long tiles = (long long)N * (H * W / TILE_PIX);
    if (tiles >= (1ll << 31) / TILE_PIX) return GP_ERR_SHAPE;
    static int sms_of[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return GP_ERR_UNSUPPORTED;
    if (!sms_of[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_gn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_gn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_gn_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair::P_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_gn_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair::P_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_gn_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair::P_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_gn_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair::P_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
    }
    const int sms = sms_of[dev];
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap mx, mw, my;
    if (!make_out_map(&my, y, tiles * TILE_PIX)) return GP_ERR_UNSUPPORTED;
    if (conv_pair_mode()) {
        // CTA pair: each CTA loads the slab of its own 128 pixels and its half of the weight rows
        if (!make_act_map(&mx, x, N, H, W, Cin, BM / W + 2) || !make_map(&mw, w_packed, BN, 9 * Cin, BN / 2)) return GP_ERR_UNSUPPORTED;
        const long long pairs = tiles < sms / 2 ? tiles : sms / 2;
        const unsigned grid = (unsigned)(2 * pairs);
        int lgW = 0;
        while ((1 << lgW) < W) ++lgW;
#define GP_PAIR_LAUNCH(ST, XF)                                                                                                               \
    conv3x3_gn_pair_kernel<ST, XF><<<grid, CONV_THREADS, pair::P_SMEM_BYTES, st>>>(mx, mw, my, partial, (int)tiles, H * W / TILE_PIX, BM / W, \
                                                                                   Cin / BK, W * BK * 2, in_stats, in_gamma, in_beta, H, lgW)
        if (partial) { if (xform) GP_PAIR_LAUNCH(true, true); else GP_PAIR_LAUNCH(true, false); }
        else { if (xform) GP_PAIR_LAUNCH(false, true); else GP_PAIR_LAUNCH(false, false); }
#undef GP_PAIR_LAUNCH
    } else {
        if (!make_act_map(&mx, x, N, H, W, Cin, TILE_PIX / W + 2) || !make_map(&mw, w_packed, BN, 9 * Cin, BN)) return GP_ERR_UNSUPPORTED;
        const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
        if (partial)
            conv3x3_gn_kernel<true><<<grid, CONV_THREADS, SMEM_BYTES, st>>>(mx, mw, my, partial, (int)tiles, H * W / TILE_PIX, BM / W, Cin / BK, W * BK * 2);
        else
            conv3x3_gn_kernel<false><<<grid, CONV_THREADS, SMEM_BYTES, st>>>(mx, mw, my, nullptr, (int)tiles, H * W / TILE_PIX, BM / W, Cin / BK, W * BK * 2);
    }
    gp::g_launches += 1;
    return (int)cudaGetLastError();
}

int gp_conv3x3_gn_bf16(const void *x, const void *w_packed, void *y, float *partial, int N, int H, int W, int Cin, int Cout, void *stream) {
    return conv3x3_impl(x, w_packed, y, partial, N, H, W, Cin, Cout, stream, nullptr, nullptr, nullptr);
}

int gp_conv3x3_gn_bf16_fused_in(const void *x_raw, const float *in_stats, const float *in_gamma, const float *in_beta, const void *w_packed,
                                void *y, float *partial, int N, int H, int W, int Cin, int Cout, void *stream) {
    if (!in_stats) return GP_ERR_NULL;
    return conv3x3_impl(x_raw, w_packed, y, partial, N, H, W, Cin, Cout, stream, in_stats, in_gamma, in_beta);
}

}  // extern "C"
