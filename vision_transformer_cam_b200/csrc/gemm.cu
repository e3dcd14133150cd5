// tcgen05 / TMEM / TMA GEMM for the dense contractions of the ViT forward (SURVEY K1,K3,K5,K6,K7):
//   out[M,N] = A[M,K] . W[N,K]^T + bias  (+ GELU | + residual | patch-embed row remap + pos_embed)
// A and W are bf16, K-major (row-major [rows,K]); accumulation is fp32 in TMEM.
//
// Kernel shape (one persistent CTA per SM, 10 warps):
//   warp 0      TMA producer: 4-stage ring of {A 128x64, W 256x64} tiles, 128-byte swizzle, mbarrier tx-count
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=16 per instruction)
//   warps 2..9  epilogue: tcgen05.ld 32 lanes x 32 columns, bias/GELU/residual in registers, vector stores.
//               Two TMEM accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cstdlib>

#include "common.cuh"
#include "ops.h"
#include "tma_host.h"

namespace vtc {

namespace gemm {
constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;   // + barriers + alignment slack

struct Params {
    const float* bias;
    const float* residual;
    const float* pos;
    void* out;
    int M, N, K;
    int tokens;    // patch-embed epilogue: tokens per image (P = tokens - 1)
    int split;     // operands are (hi | lo) bf16 halves [rows, 2K]: three K segments hi.hi, lo.hi, hi.lo
};
}  // namespace gemm

template <int EPI>
__global__ void __launch_bounds__(gemm::THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const gemm::Params p) {
    using namespace gemm;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_m = (p.M + BM - 1) / BM;
    const int num_n = p.N / BN;
    const int num_tiles = num_m * num_n;
    const int nk1 = p.K / BK;
    const int num_k = p.split ? 3 * nk1 : nk1;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / num_n) * BM;
                const int n0 = (tile % num_n) * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    const int seg = kb / nk1, kr = (kb - seg * nk1) * BK;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    uint8_t* a_dst = smem + s * STAGE_BYTES;
                    tma_load_2d(a_dst, &tmA, &full_bar[s], kr + (seg == 1 ? p.K : 0), m0);
                    tma_load_2d(a_dst + A_BYTES, &tmB, &full_bar[s], kr + (seg == 2 ? p.K : 0), n0);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 1024, 16);
                        const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 1024, 16);
                        umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        const int ew = warp - 2;             // 0..7
        const int quarter = warp & 3;        // TMEM lane quarter this warp may access
        const int half = ew >> 2;            // which 128-column half of the accumulator
        int as = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / num_n) * BM;
            const int n0 = (tile % num_n) * BN;
            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            const int row = m0 + quarter * 32 + lane;
            const bool row_ok = row < p.M;
            size_t out_row = static_cast<size_t>(row);
            const float* pos_row = nullptr;
            if (EPI == VTC_EPI_PATCH_EMBED) {
                const int P = p.tokens - 1;
                const int b = row / P;
                const int pp = row - b * P;
                out_row = static_cast<size_t>(b) * p.tokens + 1 + pp;
                pos_row = p.pos + static_cast<size_t>(1 + pp) * p.N;
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = half * 128 + c * 32;
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + col0, r);
                tmem_ld_wait();
                const int n = n0 + col0;
                const float4* bias4 = reinterpret_cast<const float4*>(p.bias + n);
                if (EPI == VTC_EPI_BIAS || EPI == VTC_EPI_BIAS_GELU) {
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = __ldg(bias4 + j);
                        float v0 = __uint_as_float(r[4 * j + 0]) + b4.x;
                        float v1 = __uint_as_float(r[4 * j + 1]) + b4.y;
                        float v2 = __uint_as_float(r[4 * j + 2]) + b4.z;
                        float v3 = __uint_as_float(r[4 * j + 3]) + b4.w;
                        if (EPI == VTC_EPI_BIAS_GELU) {
                            v0 = gelu_erf_mufu(v0); v1 = gelu_erf_mufu(v1); v2 = gelu_erf_mufu(v2); v3 = gelu_erf_mufu(v3);
                        }
                        o[2 * j] = pack_bf16x2(v0, v1);
                        o[2 * j + 1] = pack_bf16x2(v2, v3);
                    }
                    if (row_ok) {
                        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.N + n;
#pragma unroll
                        for (int j = 0; j < 4; ++j) st_u4(dst + 8 * j, make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
                    }
                } else {
                    if (row_ok) {
                        float* dst = reinterpret_cast<float*>(p.out) + out_row * p.N + n;
                        const float* add = (EPI == VTC_EPI_BIAS_RESIDUAL) ? p.residual + out_row * p.N + n : pos_row + n;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(bias4 + j);
                            const float4 a4 = *reinterpret_cast<const float4*>(add + 4 * j);
                            float4 v;
                            v.x = __uint_as_float(r[4 * j + 0]) + b4.x + a4.x;
                            v.y = __uint_as_float(r[4 * j + 1]) + b4.y + a4.y;
                            v.z = __uint_as_float(r[4 * j + 2]) + b4.z + a4.z;
                            v.w = __uint_as_float(r[4 * j + 3]) + b4.w + a4.w;
                            st_f4(dst + 4 * j, v);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const gemm::Params& p, cudaStream_t stream) {
    static SmemOptIn optin;
    int rc_ = optin.ensure(reinterpret_cast<const void*>(gemm_bf16_kernel<EPI>), gemm::SMEM_BYTES);
    if (rc_ != VTC_OK) return rc_;
    const int tiles = cdiv(p.M, gemm::BM) * (p.N / gemm::BN);
    const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
    gemm_bf16_kernel<EPI><<<grid, gemm::THREADS, gemm::SMEM_BYTES, stream>>>(tmA, tmB, p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ======================================================================================================================
// v2: CTA-pair kernel (tcgen05 cta_group::2).  One 256x256 output tile per pair: each CTA stages its own 128 rows of A
// and 128 of the 256 W rows per k-block (32 KB instead of 48 KB per CTA and k-block -> 1.5x less L2->SM traffic, which
// is what bounds the single-CTA kernel), the leader issues M=256 MMAs that write 128 accumulator rows into each CTA's
// TMEM, and each CTA drains its own half through a shared-memory staged, fully coalesced TMA store:
//   bf16 outputs           cp.async.bulk.tensor store of 128x64 boxes
//   fp32 residual stream   cp.reduce.async.bulk.tensor .add of 128x32 boxes: out += acc + bias happens in L2, the SMs
//                          never read the residual.
namespace gemm2 {
// Ablation switches for timing experiments (libraries built with -DVTC_GEMM_ABLATE=k, tools/build_ablate.sh gK; results WRONG by
// construction): bit 0 = no GELU arithmetic in the fc1 epilogue, bit 1 = no staging + bulk store of the bf16 output tiles,
// bit 2 = no tcgen05.ld of the accumulator in the bf16-output epilogues.  0 in every shipped build.  Measured (B = 256, us per
// launch): qkv 141 whatever is removed (the epilogue is hidden: operand-feed / MMA bound); fc1 195-206 -> 182 without the GELU
// arithmetic, 176 with the whole epilogue removed.  Sixteen instead of eight epilogue warps for fc1 (one 64-column group per
// warp, 96 registers) did NOT recover that: 205-213 us.
#ifndef VTC_GEMM_ABLATE
#define VTC_GEMM_ABLATE 0
#endif
constexpr int ABLATE = VTC_GEMM_ABLATE;
constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;     // 16 KB: this CTA's half of the W tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int OUT_BUF_BYTES = 128 * 128;       // 128 rows x 128 B staging tile
constexpr int EPI_RESID_LN = 4;                // internal epilogue id (see below); the public ids are VTC_EPI_*
// Patch embedding (vit_model.py:64-83, 308-314): out[b, 1 + p, :] = acc + bias + pos_embed[1 + p, :] in fp32.  The row remap
// (196 patch rows per image land behind that image's CLS row) is done by the TMA: both the patch matrix (A) and the output are
// 3-D maps [cols, patches, B] (the output's base one row into the token buffer), and a 256-row pair tile never leaves its
// image -- tile (b, q) covers patches q*256 .. of image b, the rows past the image's last patch are zero-filled on the load
// and clipped on the store (196 patches: one tile per image, 23 % of its MMA rows idle; a TMA store does not take the
// negative start coordinate that a tile crossing into the next image would need).
constexpr int EPI_PATCH = 5;
__host__ __device__ constexpr int stages_of(int) { return 5; }
__host__ __device__ constexpr int out_bufs_of(int) { return 4; }       // [2 halves][2 buffers]
__host__ __device__ constexpr int off_out_of(int epi) { return stages_of(epi) * STAGE_BYTES; }
__host__ __device__ constexpr int off_bar_of(int epi) { return off_out_of(epi) + out_bufs_of(epi) * OUT_BUF_BYTES; }
__host__ __device__ constexpr int smem_bytes_of(int epi) { return off_bar_of(epi) + 256; }
static_assert(smem_bytes_of(0) <= 232448 && smem_bytes_of(EPI_RESID_LN) <= 232448, "gemm2 smem budget");

struct Params {
    const float* bias;     // [N]; LayerNorm-folded GEMMs: c[n] = bias[n] + sum_k beta[k] W[n,k]
    int M, N, K;
    int split;     // see gemm::Params::split; bf16 outputs are then written as (hi | lo) halves [M, 2N]
    int reverse;   // walk the tiles from the last row block to the first (see model.cu: alternating sweep direction)
    // LayerNorm fusion (bf16 mode).  Row statistics travel as partial (sum, sum of squares) pairs per 128-column slice of
    // the residual stream: stats[row][slice][2], `nslices` = D / 128.
    const float* g;        // folded GEMM: g[n] = sum_k W'[n,k] with W' = bf16(gamma * W)
    float* stats;          // folded GEMM: read; residual epilogue: written
    int nslices;
    float eps;
    __nv_bfloat16* out_bf16;   // residual epilogue: bf16 copy of the new residual stream [M,N]
    const float* pos;          // EPI_PATCH: pos_embed [1 + patches, N]
    int patches;               // EPI_PATCH: patch rows per image
    int tiles_per_img;         // EPI_PATCH: 256-row pair tiles per image
};
}  // namespace gemm2

// LayerNorm fusion (bf16 mode): the separate LayerNorm kernel (a full read of the fp32 residual stream + a bf16 write per
// call, 24 calls per forward) is folded into the GEMMs either side of it.
//   EPI_RESID_LN (proj, fc2): t_new = residual + A.W^T + bias is formed in registers -- the fp32 residual chunk is fetched by
//     TMA into the staging tile while the accumulator is still being computed -- and leaves the SM three ways: fp32 t_new
//     (TMA store), bf16(t_new) (TMA store: the A operand of the next GEMM) and per-row partial sums (sum, sum of squares)
//     of this CTA's 128 columns.
//   FOLD (qkv, fc1): LN(t).W^T = rstd (t.W'^T - mean g) + c with W' = gamma * W, g[n] = sum_k W'[n,k],
//     c[n] = bias[n] + sum_k beta[k] W[n,k]: the GEMM runs on bf16(t) and the epilogue applies the two row scalars.
template <int EPI, bool SPLIT, bool FOLD>
__global__ void __launch_bounds__(gemm2::THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ CUtensorMap tmH, const gemm2::Params p) {
    using namespace gemm2;
    constexpr int STAGES = stages_of(EPI);
    constexpr int OFF_OUT = off_out_of(EPI);
    constexpr int OFF_BAR = off_bar_of(EPI);
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // used in the leader CTA only
    uint64_t* empty_bar = full_bar + STAGES;                             // per CTA
    uint64_t* tfull_bar = empty_bar + STAGES;                            // per CTA
    uint64_t* tempty_bar = tfull_bar + 2;                                // used in the leader CTA only
    uint64_t* res_bar = tempty_bar + 2;                                  // [2 halves][2 buffers] residual chunk landed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        if (EPI == EPI_RESID_LN) tma_prefetch_desc(&tmR);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 2 * EPI_WARPS);     // epilogue warps of BOTH CTAs
        }
        for (int s = 0; s < 4; ++s) mbar_init(&res_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_m = (EPI == EPI_PATCH) ? (p.M / p.patches) * p.tiles_per_img : (p.M + 2 * BM - 1) / (2 * BM);
    const int num_n = p.N / BN;
    const int num_tiles = num_m * num_n;
    const int nk1 = p.K / BK;
    const int num_k = SPLIT ? 3 * nk1 : nk1;
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t_ = pair; t_ < num_tiles; t_ += npairs) {
                const int tile = p.reverse ? num_tiles - 1 - t_ : t_;
                const int m0 = (tile / num_n) * (2 * BM) + rank * BM;
                const int n0 = (tile % num_n) * BN + rank * (BN / 2);
                const int pimg = (EPI == EPI_PATCH) ? (tile / num_n) / p.tiles_per_img : 0;                       // EPI_PATCH: image,
                const int prow = (EPI == EPI_PATCH) ? ((tile / num_n) % p.tiles_per_img) * (2 * BM) + rank * BM : 0;   // first patch of this CTA
                for (int kb = 0; kb < num_k; ++kb) {
                    const int seg = kb / nk1, kr = (kb - seg * nk1) * BK;      // split: A = hi, lo, hi ; W = hi, hi, lo
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);     // both CTAs' bytes land on this barrier
                    uint8_t* a_dst = smem + s * STAGE_BYTES;
                    if (EPI == EPI_PATCH) tma_load_3d_2sm(a_dst, &tmA, &full_bar[s], kr + (seg == 1 ? p.K : 0), prow, pimg);
                    else tma_load_2d_2sm(a_dst, &tmA, &full_bar[s], kr + (seg == 1 ? p.K : 0), m0);
                    tma_load_2d_2sm(a_dst + A_BYTES, &tmB, &full_bar[s], kr + (seg == 2 ? p.K : 0), n0);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int tile = pair; tile < num_tiles; tile += npairs) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 1024, 16);
                        const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 1024, 16);
                        umma_bf16_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty_bar[s], 3);          // frees the stage in both CTAs
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_2sm(&tfull_bar[as], 3);             // accumulator ready in both CTAs
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const int r_local = quarter * 32 + lane;
        const bool issuer = (quarter == 0) && (lane == 0);
        const uint32_t bar_id = 1 + half;
        uint8_t* stage_out = smem + OFF_OUT + half * (out_bufs_of(EPI) / 2) * OUT_BUF_BYTES;
        constexpr int CHUNK_COLS = (EPI == VTC_EPI_BIAS_RESIDUAL || EPI == EPI_RESID_LN || EPI == EPI_PATCH) ? 32 : 64;     // 128-byte rows
        constexpr int NCHUNK = 128 / CHUNK_COLS;
        int as = 0;
        uint32_t aph = 0;
        int buf = 0;
        uint32_t res_ph[2] = {0, 0};      // EPI_RESID_LN: phases of this half's two residual buffers
        for (int t_ = pair; t_ < num_tiles; t_ += npairs) {
            const int tile = p.reverse ? num_tiles - 1 - t_ : t_;
            const int m0 = (tile / num_n) * (2 * BM) + rank * BM;
            const int n0 = (tile % num_n) * BN;
            if constexpr (EPI == EPI_RESID_LN) {
                // Residual chunks 0 and 1 of this tile are fetched while the tensor core is still accumulating.  The staging tiles
                // are free once the bulk stores issued from them for the previous tile have been read.
                if (issuer) {
                    tma_store_wait_read<0>();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        mbar_arrive_expect_tx(&res_bar[half * 2 + c], OUT_BUF_BYTES);
                        tma_load_2d(stage_out + c * OUT_BUF_BYTES, &tmR, &res_bar[half * 2 + c], n0 + half * 128 + c * 32, m0);
                    }
                    // pull the residual of the NEXT tile of this CTA into L2 now: only two chunks fit in shared memory, and an HBM
                    // round trip per chunk would make this epilogue latency-bound
                    if (t_ + npairs < num_tiles) {
                        const int nt = p.reverse ? num_tiles - 1 - (t_ + npairs) : t_ + npairs;
                        const int nm0 = (nt / num_n) * (2 * BM) + rank * BM, nn0 = (nt % num_n) * BN + half * 128;
#pragma unroll
                        for (int c = 0; c < 4; ++c) tma_prefetch_2d(&tmR, nn0 + c * 32, nm0);
                    }
                }
            }
            // row scalars of the folded LayerNorm: mean and rstd from the partial sums of the producer GEMM (fetched while the
            // tensor core is still accumulating this tile)
            float ln_rs = 1.f, ln_nrm = 0.f;      // rstd and -rstd * mean
            if constexpr (FOLD) {
                const int row = m0 + r_local;
                if (row < p.M) {
                    const float2* st2 = reinterpret_cast<const float2*>(p.stats) + static_cast<size_t>(row) * p.nslices;
                    float a = 0.f, b = 0.f;
                    for (int i = 0; i < p.nslices; ++i) {
                        const float2 v = __ldg(st2 + i);
                        a += v.x;
                        b += v.y;
                    }
                    const float invd = 1.0f / static_cast<float>(p.K);
                    const float mean = a * invd;
                    const float var = fmaxf(fmaf(-mean, mean, b * invd), 0.f);
                    ln_rs = rsqrtf(var + p.eps);
                    ln_nrm = -ln_rs * mean;
                }
            }
            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * 128;
            if constexpr (EPI == EPI_RESID_LN) {
                const int row = m0 + r_local;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int pr = 0; pr < 2; ++pr) {
#pragma unroll
                for (int fb = 0; fb < 2; ++fb) {                   // fb = staging buffer of chunk c (static: hb stays in registers)
                    const int c = pr * 2 + fb;
                    const int col0 = half * 128 + c * 32;
                    uint8_t* frow = stage_out + fb * OUT_BUF_BYTES + r_local * 128;
                    uint32_t acc[32];
                    uint32_t hb[16];                              // bf16(t_new) of this chunk
                    tmem_ld_32x32b_x32(t_row + c * 32, acc);
                    mbar_wait(&res_bar[half * 2 + fb], res_ph[fb]);
                    res_ph[fb] ^= 1;
                    tmem_ld_wait();
                    if (c == 3) {                                 // the accumulator is in registers: hand the TMEM stage back early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], 0);
                    }
                    const float4* bias4 = reinterpret_cast<const float4*>(p.bias + n0 + col0);
#pragma unroll
                    for (int g8 = 0; g8 < 8; ++g8) {
                        float4* slot = reinterpret_cast<float4*>(frow + ((g8 ^ (r_local & 7)) * 16));
                        const float4 r4 = *slot;
                        const float4 b4 = __ldg(bias4 + g8);
                        float4 v;
                        v.x = (__uint_as_float(acc[4 * g8 + 0]) + b4.x) + r4.x;
                        v.y = (__uint_as_float(acc[4 * g8 + 1]) + b4.y) + r4.y;
                        v.z = (__uint_as_float(acc[4 * g8 + 2]) + b4.z) + r4.z;
                        v.w = (__uint_as_float(acc[4 * g8 + 3]) + b4.w) + r4.w;
                        *slot = v;
                        s1 += (v.x + v.y) + (v.z + v.w);
                        s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
                        hb[2 * g8] = pack_bf16x2(v.x, v.y);
                        hb[2 * g8 + 1] = pack_bf16x2(v.z, v.w);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (issuer) {
                        tma_store_2d(&tmO, stage_out + fb * OUT_BUF_BYTES, n0 + col0, m0);
                        tma_store_commit();
                    }
                    if (row < p.M) {
                        // bf16(t_new): 64 contiguous bytes per row straight from registers (the fp32 copy goes through the TMA store)
                        __nv_bfloat16* hdst = p.out_bf16 + static_cast<size_t>(row) * p.N + n0 + col0;
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) st_u4(hdst + 8 * g4, make_uint4(hb[4 * g4], hb[4 * g4 + 1], hb[4 * g4 + 2], hb[4 * g4 + 3]));
                    }
                    if (c < 2 && issuer) {
                        // refill this fp32 buffer with chunk c+2 once its store has been read
                        tma_store_wait_read<0>();
                        mbar_arrive_expect_tx(&res_bar[half * 2 + fb], OUT_BUF_BYTES);
                        tma_load_2d(stage_out + fb * OUT_BUF_BYTES, &tmR, &res_bar[half * 2 + fb], n0 + half * 128 + (c + 2) * 32, m0);
                    }
                }
                }
                if (row < p.M) {
                    float* dst = p.stats + (static_cast<size_t>(row) * p.nslices + (n0 >> 7) + half) * 2;
                    *reinterpret_cast<float2*>(dst) = make_float2(s1, s2);
                }
                if (++as == 2) { as = 0; aph ^= 1; }
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                const int col0 = half * 128 + c * CHUNK_COLS;
                const float4* bias4 = reinterpret_cast<const float4*>(p.bias + n0 + col0);
                uint32_t pk[32];
                uint32_t pl[(EPI == VTC_EPI_BIAS_RESIDUAL || EPI == EPI_PATCH || !SPLIT) ? 1 : 32];      // split mode: low halves
                if (EPI == VTC_EPI_BIAS_RESIDUAL || EPI == EPI_PATCH) {
                    tmem_ld_32x32b_x32(t_row + c * 32, pk);
                    tmem_ld_wait();
                    // EPI_PATCH: + pos_embed[1 + patch index of my row] (what the reference adds after the concat, vit_model.py:313)
                    const int prow = (EPI == EPI_PATCH) ? ((tile / num_n) % p.tiles_per_img) * (2 * BM) + rank * BM + r_local : 0;
                    const float4* pos4 = (EPI == EPI_PATCH)
                        ? reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + (prow < p.patches ? prow : p.patches - 1)) * p.N + n0 + col0) : nullptr;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 b4 = __ldg(bias4 + j);
                        if (EPI == EPI_PATCH) {
                            const float4 q4 = __ldg(pos4 + j);
                            b4.x += q4.x; b4.y += q4.y; b4.z += q4.z; b4.w += q4.w;
                        }
                        pk[4 * j + 0] = __float_as_uint(__uint_as_float(pk[4 * j + 0]) + b4.x);
                        pk[4 * j + 1] = __float_as_uint(__uint_as_float(pk[4 * j + 1]) + b4.y);
                        pk[4 * j + 2] = __float_as_uint(__uint_as_float(pk[4 * j + 2]) + b4.z);
                        pk[4 * j + 3] = __float_as_uint(__uint_as_float(pk[4 * j + 3]) + b4.w);
                    }
                } else {
                    uint32_t r0[32], r1[32];
                    if constexpr ((ABLATE & 4) != 0) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) { r0[q] = __float_as_uint(static_cast<float>(q + lane)); r1[q] = r0[q]; }
                    } else {
                        tmem_ld_32x32b_x32(t_row + c * 64, r0);
                        tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = __ldg(bias4 + j);
                        const float4 c4 = __ldg(bias4 + 8 + j);
                        float v0, v1, v2, v3, w0, w1, w2, w3;
                        if constexpr (FOLD) {
                            // rstd * acc + (-rstd * mean) * g[n] + c[n]
                            const float4* g4 = reinterpret_cast<const float4*>(p.g + n0 + col0);
                            const float4 ga = __ldg(g4 + j), gb = __ldg(g4 + 8 + j);
                            const uint64_t rs2 = pack2(ln_rs, ln_rs), nm2 = pack2(ln_nrm, ln_nrm);
                            unpack2(fma2(pack2u(r0[4 * j + 0], r0[4 * j + 1]), rs2, fma2(nm2, pack2(ga.x, ga.y), pack2(b4.x, b4.y))), v0, v1);
                            unpack2(fma2(pack2u(r0[4 * j + 2], r0[4 * j + 3]), rs2, fma2(nm2, pack2(ga.z, ga.w), pack2(b4.z, b4.w))), v2, v3);
                            unpack2(fma2(pack2u(r1[4 * j + 0], r1[4 * j + 1]), rs2, fma2(nm2, pack2(gb.x, gb.y), pack2(c4.x, c4.y))), w0, w1);
                            unpack2(fma2(pack2u(r1[4 * j + 2], r1[4 * j + 3]), rs2, fma2(nm2, pack2(gb.z, gb.w), pack2(c4.z, c4.w))), w2, w3);
                        } else {
                            unpack2(add2(pack2u(r0[4 * j + 0], r0[4 * j + 1]), pack2(b4.x, b4.y)), v0, v1);
                            unpack2(add2(pack2u(r0[4 * j + 2], r0[4 * j + 3]), pack2(b4.z, b4.w)), v2, v3);
                            unpack2(add2(pack2u(r1[4 * j + 0], r1[4 * j + 1]), pack2(c4.x, c4.y)), w0, w1);
                            unpack2(add2(pack2u(r1[4 * j + 2], r1[4 * j + 3]), pack2(c4.z, c4.w)), w2, w3);
                        }
                        if (EPI == VTC_EPI_BIAS_GELU && (ABLATE & 1) == 0) {
                            if (SPLIT) {      // fp32 mode: exact erf instead of the 4e-7 polynomial
                                v0 = gelu_erf_exact(v0); v1 = gelu_erf_exact(v1); v2 = gelu_erf_exact(v2); v3 = gelu_erf_exact(v3);
                                w0 = gelu_erf_exact(w0); w1 = gelu_erf_exact(w1); w2 = gelu_erf_exact(w2); w3 = gelu_erf_exact(w3);
                            } else {
                                gelu2(v0, v1, v0, v1);
                                gelu2(v2, v3, v2, v3);
                                gelu2(w0, w1, w0, w1);
                                gelu2(w2, w3, w2, w3);
                            }
                        }
                        pk[2 * j] = pack_bf16x2(v0, v1);
                        pk[2 * j + 1] = pack_bf16x2(v2, v3);
                        pk[16 + 2 * j] = pack_bf16x2(w0, w1);
                        pk[16 + 2 * j + 1] = pack_bf16x2(w2, w3);
                        if (EPI != VTC_EPI_BIAS_RESIDUAL && EPI != EPI_PATCH) {
                            if (SPLIT) {
                                pl[2 * j] = pack_bf16x2(v0 - __uint_as_float(pk[2 * j] << 16), v1 - __uint_as_float(pk[2 * j] & 0xffff0000u));
                                pl[2 * j + 1] = pack_bf16x2(v2 - __uint_as_float(pk[2 * j + 1] << 16), v3 - __uint_as_float(pk[2 * j + 1] & 0xffff0000u));
                                pl[16 + 2 * j] = pack_bf16x2(w0 - __uint_as_float(pk[16 + 2 * j] << 16), w1 - __uint_as_float(pk[16 + 2 * j] & 0xffff0000u));
                                pl[16 + 2 * j + 1] = pack_bf16x2(w2 - __uint_as_float(pk[16 + 2 * j + 1] << 16), w3 - __uint_as_float(pk[16 + 2 * j + 1] & 0xffff0000u));
                            }
                        }
                    }
                }
                // Stage one 128 x 128-byte tile in shared memory (swizzled) and hand it to the TMA; staging buffer `buf` is free
                // once the bulk store issued from it two tiles ago has read it.
                auto stage_and_store = [&](const uint32_t* v, int col) {
                    if (issuer) tma_store_wait_read<1>();
                    named_bar_sync(bar_id, 128);
                    uint8_t* row_ptr = stage_out + buf * OUT_BUF_BYTES + r_local * 128;
#pragma unroll
                    for (int g8 = 0; g8 < 8; ++g8)
                        st_u4(row_ptr + ((g8 ^ (r_local & 7)) * 16), make_uint4(v[4 * g8], v[4 * g8 + 1], v[4 * g8 + 2], v[4 * g8 + 3]));
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (issuer) {
                        if (EPI == VTC_EPI_BIAS_RESIDUAL) {
                            tma_reduce_add_2d(&tmO, stage_out + buf * OUT_BUF_BYTES, col, m0);
                        } else if (EPI == EPI_PATCH) {
                            const int b0 = (tile / num_n) / p.tiles_per_img, p0 = ((tile / num_n) % p.tiles_per_img) * (2 * BM) + rank * BM;
                            tma_store_3d(&tmO, stage_out + buf * OUT_BUF_BYTES, col, p0, b0);
                        } else {
                            tma_store_2d(&tmO, stage_out + buf * OUT_BUF_BYTES, col, m0);
                        }
                        tma_store_commit();
                    }
                    buf ^= 1;
                };
                if ((ABLATE & 2) == 0 || EPI == VTC_EPI_BIAS_RESIDUAL || EPI == EPI_PATCH) stage_and_store(pk, n0 + col0);
                else if (pk[lane & 31] == 0x12345678u) stage_and_store(pk, n0 + col0);      // keep the values alive
                if (EPI != VTC_EPI_BIAS_RESIDUAL && EPI != EPI_PATCH) {
                    if (SPLIT) stage_and_store(pl, p.N + n0 + col0);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], 0);      // the leader's MMA thread owns the TMEM schedule
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        if (issuer) tma_store_wait_all<0>();
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();              // the peer may still read this CTA's smem / signal its barriers until here
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

template <int EPI, bool SPLIT, bool FOLD = false>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const gemm2::Params& p, cudaStream_t stream,
                        const CUtensorMap* tmR = nullptr, const CUtensorMap* tmH = nullptr) {
    static SmemOptIn optin;
    int rc_ = optin.ensure(reinterpret_cast<const void*>(gemm2_bf16_kernel<EPI, SPLIT, FOLD>), gemm2::smem_bytes_of(EPI));
    if (rc_ != VTC_OK) return rc_;
    const int tiles = (EPI == gemm2::EPI_PATCH ? (p.M / p.patches) * p.tiles_per_img : cdiv(p.M, 2 * gemm2::BM)) * (p.N / gemm2::BN);
    int pairs = device_sm_count() / 2;
    if (tiles < pairs) pairs = tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(gemm2::THREADS);
    cfg.dynamicSmemBytes = gemm2::smem_bytes_of(EPI);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    note_launch();
    VTC_CUDA(cudaLaunchKernelEx(&cfg, gemm2_bf16_kernel<EPI, SPLIT, FOLD>, tmA, tmB, tmO, tmR ? *tmR : tmO, tmH ? *tmH : tmO, p));
    return VTC_OK;
}

static bool use_v1() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VTC_GEMM_V1");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

int gemm_bf16(const void* A, const void* W, const float* bias, const float* residual, const float* pos, void* out, int M,
              int N, int K, int epilogue, int tokens, cudaStream_t stream, int split, int reverse) {
    VTC_REQUIRE(A && W && bias && out, VTC_ERR_ARG, "gemm: null pointer");
    VTC_REQUIRE(M > 0 && N > 0 && K > 0, VTC_ERR_SHAPE, "gemm: empty problem %dx%dx%d", M, N, K);
    VTC_REQUIRE(K % gemm::BK == 0, VTC_ERR_SHAPE, "gemm: K=%d must be a multiple of %d", K, gemm::BK);
    VTC_REQUIRE(N % gemm::BN == 0, VTC_ERR_SHAPE, "gemm: N=%d must be a multiple of %d", N, gemm::BN);
    VTC_REQUIRE(epilogue >= VTC_EPI_BIAS && epilogue <= VTC_EPI_PATCH_EMBED, VTC_ERR_ARG, "gemm: unknown epilogue %d", epilogue);
    VTC_REQUIRE(epilogue != VTC_EPI_BIAS_RESIDUAL || residual, VTC_ERR_ARG, "gemm: residual epilogue without residual");
    VTC_REQUIRE(epilogue != VTC_EPI_PATCH_EMBED || (pos && tokens > 1 && M % (tokens - 1) == 0), VTC_ERR_ARG,
                "gemm: patch-embed epilogue needs pos_embed and M %% (tokens-1) == 0");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    // patch embedding through the CTA-pair kernel (3-D TMA store does the row remap) unless the patch count per image is
    // smaller than a 128-row staging tile (a tile would then span three images: ViT-B/32 has 49) -> single-CTA kernel
    const bool patch_pair = epilogue == VTC_EPI_PATCH_EMBED && tokens - 1 >= gemm2::BM && !use_v1();
    const bool v1 = (use_v1() && !split) || (epilogue == VTC_EPI_PATCH_EMBED && !patch_pair);
    const uint64_t kcols = static_cast<uint64_t>(split ? 2 : 1) * K;       // (hi | lo) halves side by side
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {kcols, (uint64_t)M};
        uint64_t strides[1] = {kcols * 2};
        uint32_t box[2] = {gemm::BK, 128};
        rc = make_tmap_bf16(&tmA, A, 2, dims, strides, box);
        if (rc != VTC_OK) return rc;
    }
    {
        uint64_t dims[2] = {kcols, (uint64_t)N};
        uint64_t strides[1] = {kcols * 2};
        uint32_t box[2] = {gemm::BK, v1 ? 256u : 128u};
        rc = make_tmap_bf16(&tmB, W, 2, dims, strides, box);
        if (rc != VTC_OK) return rc;
    }
    if (v1) {
        gemm::Params p{bias, residual, pos, out, M, N, K, tokens, split};
        switch (epilogue) {
            case VTC_EPI_BIAS: return launch_gemm<VTC_EPI_BIAS>(tmA, tmB, p, stream);
            case VTC_EPI_BIAS_GELU: return launch_gemm<VTC_EPI_BIAS_GELU>(tmA, tmB, p, stream);
            case VTC_EPI_BIAS_RESIDUAL: return launch_gemm<VTC_EPI_BIAS_RESIDUAL>(tmA, tmB, p, stream);
            default: return launch_gemm<VTC_EPI_PATCH_EMBED>(tmA, tmB, p, stream);
        }
    }
    CUtensorMap tmO;
    gemm2::Params p2{bias, M, N, K, split, reverse, nullptr, nullptr, 0, 0.f, nullptr, nullptr, 0, 0};
    if (patch_pair) {
        const int patches = tokens - 1;
        p2.pos = pos;
        p2.patches = patches;
        p2.tiles_per_img = cdiv(patches, 2 * gemm2::BM);
        {   // A as [K, patches, B]: rows past an image's last patch come back as zeros
            uint64_t adims[3] = {kcols, (uint64_t)patches, (uint64_t)(M / patches)};
            uint64_t astrides[2] = {kcols * 2, (uint64_t)patches * kcols * 2};
            uint32_t abox[3] = {gemm::BK, 128, 1};
            rc = make_tmap_bf16(&tmA, A, 3, adims, astrides, abox);
            if (rc != VTC_OK) return rc;
        }
        uint64_t dims[3] = {(uint64_t)N, (uint64_t)patches, (uint64_t)(M / patches)};
        uint64_t strides[2] = {(uint64_t)N * 4, (uint64_t)tokens * N * 4};
        uint32_t box[3] = {32, 128, 1};
        rc = make_tmap_f32(&tmO, static_cast<float*>(out) + N, 3, dims, strides, box);       // base: row 1 of image 0 (behind its CLS row)
        if (rc != VTC_OK) return rc;
        return split ? launch_gemm2<gemm2::EPI_PATCH, true>(tmA, tmB, tmO, p2, stream) : launch_gemm2<gemm2::EPI_PATCH, false>(tmA, tmB, tmO, p2, stream);
    }
    if (epilogue == VTC_EPI_BIAS_RESIDUAL) {
        // out = residual + A.W^T + bias, with the add done by the TMA reduction into `out`
        if (out != static_cast<const void*>(residual))
            VTC_CUDA(cudaMemcpyAsync(out, residual, static_cast<size_t>(M) * N * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
        uint64_t strides[1] = {(uint64_t)N * 4};
        uint32_t box[2] = {32, 128};
        rc = make_tmap_f32(&tmO, out, 2, dims, strides, box);
        if (rc != VTC_OK) return rc;
        return split ? launch_gemm2<VTC_EPI_BIAS_RESIDUAL, true>(tmA, tmB, tmO, p2, stream)
                     : launch_gemm2<VTC_EPI_BIAS_RESIDUAL, false>(tmA, tmB, tmO, p2, stream);
    }
    const uint64_t ocols = static_cast<uint64_t>(split ? 2 : 1) * N;
    uint64_t dims[2] = {ocols, (uint64_t)M};
    uint64_t strides[1] = {ocols * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&tmO, out, 2, dims, strides, box);
    if (rc != VTC_OK) return rc;
    if (epilogue == VTC_EPI_BIAS)
        return split ? launch_gemm2<VTC_EPI_BIAS, true>(tmA, tmB, tmO, p2, stream) : launch_gemm2<VTC_EPI_BIAS, false>(tmA, tmB, tmO, p2, stream);
    return split ? launch_gemm2<VTC_EPI_BIAS_GELU, true>(tmA, tmB, tmO, p2, stream) : launch_gemm2<VTC_EPI_BIAS_GELU, false>(tmA, tmB, tmO, p2, stream);
}

// ---- LayerNorm-fused GEMMs (bf16 mode) -------------------------------------------------------------------------------------
static int make_ab_maps(CUtensorMap* tmA, CUtensorMap* tmB, const void* A, const void* W, int M, int N, int K) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {gemm::BK, 128};
    int rc = make_tmap_bf16(tmA, A, 2, dims, strides, box);
    if (rc != VTC_OK) return rc;
    uint64_t dimsb[2] = {(uint64_t)K, (uint64_t)N};
    return make_tmap_bf16(tmB, W, 2, dimsb, strides, box);
}

int gemm_resid_ln(const void* A, const void* W, const float* bias, const float* residual, float* out, void* out_bf16, float* stats, int M, int N,
                  int K, cudaStream_t stream, int reverse) {
    VTC_REQUIRE(A && W && bias && residual && out && out_bf16 && stats, VTC_ERR_ARG, "gemm_resid_ln: null pointer");
    VTC_REQUIRE(M > 0 && N > 0 && K > 0 && K % gemm::BK == 0 && N % gemm::BN == 0, VTC_ERR_SHAPE, "gemm_resid_ln: unsupported %dx%dx%d", M, N, K);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    CUtensorMap tmA, tmB, tmO, tmR;
    if ((rc = make_ab_maps(&tmA, &tmB, A, W, M, N, K)) != VTC_OK) return rc;
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)N * 4};
    uint32_t box[2] = {32, 128};
    if ((rc = make_tmap_f32(&tmO, out, 2, dims, strides, box)) != VTC_OK) return rc;
    if ((rc = make_tmap_f32(&tmR, residual, 2, dims, strides, box)) != VTC_OK) return rc;
    gemm2::Params p{bias, M, N, K, 0, reverse, nullptr, stats, N / 128, 0.f, static_cast<__nv_bfloat16*>(out_bf16)};
    return launch_gemm2<gemm2::EPI_RESID_LN, false, false>(tmA, tmB, tmO, p, stream, &tmR, nullptr);
}

int gemm_lnfold(const void* A, const void* W, const float* c, const float* g, const float* stats, float eps, void* out, int M, int N, int K,
                int gelu, cudaStream_t stream, int reverse) {
    VTC_REQUIRE(A && W && c && g && stats && out, VTC_ERR_ARG, "gemm_lnfold: null pointer");
    VTC_REQUIRE(M > 0 && N > 0 && K > 0 && K % 128 == 0 && N % gemm::BN == 0, VTC_ERR_SHAPE, "gemm_lnfold: unsupported %dx%dx%d", M, N, K);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    CUtensorMap tmA, tmB, tmO;
    if ((rc = make_ab_maps(&tmA, &tmB, A, W, M, N, K)) != VTC_OK) return rc;
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&tmO, out, 2, dims, strides, box)) != VTC_OK) return rc;
    gemm2::Params p{c, M, N, K, 0, reverse, g, const_cast<float*>(stats), K / 128, eps, nullptr};
    return gelu ? launch_gemm2<VTC_EPI_BIAS_GELU, false, true>(tmA, tmB, tmO, p, stream) : launch_gemm2<VTC_EPI_BIAS, false, true>(tmA, tmB, tmO, p, stream);
}

}  // namespace vtc

extern "C" int vtc_gemm_resid_ln(const void* A, const void* W, const float* bias, const float* residual, float* out, void* out_bf16,
                                 float* stats, int32_t M, int32_t Nout, int32_t K, void* stream) {
    return vtc::gemm_resid_ln(A, W, bias, residual, out, out_bf16, stats, M, Nout, K, static_cast<cudaStream_t>(stream), 0);
}
extern "C" int vtc_gemm_lnfold(const void* A, const void* W, const float* c, const float* g, const float* stats, float eps, void* out,
                               int32_t M, int32_t Nout, int32_t K, int32_t gelu, void* stream) {
    return vtc::gemm_lnfold(A, W, c, g, stats, eps, out, M, Nout, K, gelu, static_cast<cudaStream_t>(stream), 0);
}

extern "C" int vtc_gemm_bf16(const void* A, const void* W, const float* bias, const float* residual, const float* pos, void* out,
                             int32_t M, int32_t Nout, int32_t K, int32_t epilogue, int32_t tokens, void* stream) {
    return vtc::gemm_bf16(A, W, bias, residual, pos, out, M, Nout, K, epilogue, tokens, static_cast<cudaStream_t>(stream), 0, 0);
}
extern "C" int vtc_gemm_split(const void* A, const void* W, const float* bias, const float* residual, const float* pos, void* out,
                              int32_t M, int32_t Nout, int32_t K, int32_t epilogue, int32_t tokens, void* stream) {
    return vtc::gemm_bf16(A, W, bias, residual, pos, out, M, Nout, K, epilogue, tokens, static_cast<cudaStream_t>(stream), 1, 0);
}
