// tcgen05 / TMEM / TMA GEMM for the dense contractions of the ViT forward (SURVEY K1,K3,K5,K6,K7):
//   out[M,N] = A[M,K] . W[N,K]^T + bias  (+ GELU | + residual | patch-embed row remap + pos_embed)
// A and W are bf16, K-major (row-major [rows,K]); accumulation is fp32 in TMEM.
//
// Kernel shape (one persistent CTA per SM, 10 warps):
//   warp 0      TMA producer: 4-stage ring of {A 128x64, W 256x64} tiles, 128-byte swizzle, mbarrier tx-count
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=16 per instruction)
//   warps 2..9  epilogue: tcgen05.ld 32 lanes x 32 columns, bias/GELU/residual in registers, vector stores.
//               Two TMEM accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
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
    const int num_k = p.K / BK;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / num_n) * BM;
                const int n0 = (tile % num_n) * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    uint8_t* a_dst = smem + s * STAGE_BYTES;
                    tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
                    tma_load_2d(a_dst + A_BYTES, &tmB, &full_bar[s], kb * BK, n0);
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
                            v0 = gelu_erf_fast(v0); v1 = gelu_erf_fast(v1); v2 = gelu_erf_fast(v2); v3 = gelu_erf_fast(v3);
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
    static bool configured = false;
    if (!configured) {
        VTC_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
        configured = true;
    }
    const int tiles = cdiv(p.M, gemm::BM) * (p.N / gemm::BN);
    const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
    gemm_bf16_kernel<EPI><<<grid, gemm::THREADS, gemm::SMEM_BYTES, stream>>>(tmA, tmB, p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

int gemm_bf16(const void* A, const void* W, const float* bias, const float* residual, const float* pos, void* out, int M,
              int N, int K, int epilogue, int tokens, cudaStream_t stream) {
    VTC_REQUIRE(A && W && bias && out, VTC_ERR_ARG, "gemm: null pointer");
    VTC_REQUIRE(M > 0 && N > 0 && K > 0, VTC_ERR_SHAPE, "gemm: empty problem %dx%dx%d", M, N, K);
    VTC_REQUIRE(K % gemm::BK == 0, VTC_ERR_SHAPE, "gemm: K=%d must be a multiple of %d", K, gemm::BK);
    VTC_REQUIRE(N % gemm::BN == 0, VTC_ERR_SHAPE, "gemm: N=%d must be a multiple of %d", N, gemm::BN);
    VTC_REQUIRE(epilogue != VTC_EPI_BIAS_RESIDUAL || residual, VTC_ERR_ARG, "gemm: residual epilogue without residual");
    VTC_REQUIRE(epilogue != VTC_EPI_PATCH_EMBED || (pos && tokens > 1 && M % (tokens - 1) == 0), VTC_ERR_ARG,
                "gemm: patch-embed epilogue needs pos_embed and M %% (tokens-1) == 0");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t strides[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {gemm::BK, gemm::BM};
        rc = make_tmap_bf16(&tmA, A, 2, dims, strides, box);
        if (rc != VTC_OK) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t strides[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {gemm::BK, gemm::BN};
        rc = make_tmap_bf16(&tmB, W, 2, dims, strides, box);
        if (rc != VTC_OK) return rc;
    }
    gemm::Params p{bias, residual, pos, out, M, N, K, tokens};
    switch (epilogue) {
        case VTC_EPI_BIAS: return launch_gemm<VTC_EPI_BIAS>(tmA, tmB, p, stream);
        case VTC_EPI_BIAS_GELU: return launch_gemm<VTC_EPI_BIAS_GELU>(tmA, tmB, p, stream);
        case VTC_EPI_BIAS_RESIDUAL: return launch_gemm<VTC_EPI_BIAS_RESIDUAL>(tmA, tmB, p, stream);
        case VTC_EPI_PATCH_EMBED: return launch_gemm<VTC_EPI_PATCH_EMBED>(tmA, tmB, p, stream);
        default: set_last_error("gemm: unknown epilogue %d", epilogue); return VTC_ERR_ARG;
    }
}

}  // namespace vtc

extern "C" int vtc_gemm_bf16(const void* A, const void* W, const float* bias, const float* residual, const float* pos, void* out,
                             int32_t M, int32_t Nout, int32_t K, int32_t epilogue, int32_t tokens, void* stream) {
    return vtc::gemm_bf16(A, W, bias, residual, pos, out, M, Nout, K, epilogue, tokens, static_cast<cudaStream_t>(stream));
}
