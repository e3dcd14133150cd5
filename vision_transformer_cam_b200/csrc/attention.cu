// Fused multi-head attention for the ViT blocks (vit_model.py:113-137): per (image, head)
//   S = Q K^T * scale (+ the layer>4 background mask: -100 on background keys, for foreground query rows),
//   P = softmax(S), O = P V
// with both contractions on tcgen05 (S and O accumulate in TMEM) and the softmax in fp32 registers.  Besides O the
// kernels emit what the reference reads back from the full P tensor: the CLS query row P[b,h,0,:] (mask builder,
// top-k head, per-layer maps) and, on request, the whole P (the 6-tuple's attn_weights / the rollout's head mean).
//
// Three kernels, one dispatcher (`attention`):
//   attention_cs.cu   fast path (no full P), any sequence length: two groups x two column halves per CTA
//   attention_kv.cu   full P of long sequences (second sweep) and the split-bf16 fp32 mode
//   this file         full P for N <= 256 tokens (all keys of a head in one TMEM accumulator), and the head mean of P
#include <cstdlib>

#include "common.cuh"
#include "ops.h"
#include "tma_host.h"

namespace vtc {

static unsigned long long* g_attn_trace = nullptr;   // debug hook, see vtc_debug_set_attention_trace
unsigned long long* attention_trace_buffer() { return g_attn_trace; }

// ======================================================================================================================
// v2: one CTA per (image, head, 128-query-row tile), two CTAs resident per SM so that one CTA's TMA / prologue /
// epilogue latency hides behind the other's softmax.
//   * smem 109 KB: Q 16 KB + K 32 KB; the bf16 P tile (64 KB, K-major 128B swizzle) OVERLAYS Q and K, which are dead once
//     S = Q K^T has been committed; V 32 KB (MN-major, consumed as the TMA wrote it).  TMEM 256 columns: S, then O on top.
//   * the reference's background mask -100*min(v_i + v_j, 1) is applied BY THE TENSOR CORE: one extra K-step with
//     Q_aug[i] = [v_i == 0] and K_aug[j] = -100/scale * v_j (bf16 -800 for scale 1/8), so the softmax code has no mask
//     handling at all and background query rows stay unmasked exactly like the reference (vit_model.py:348-361).
//   * softmax per element: pass 1 FMNMX on the raw accumulator; pass 2 FFMA + MUFU.EX2 + FADD + cvt/pack; TMEM loads are
//     software-pipelined one 32-column chunk ahead.
namespace attn2 {
constexpr int HD = 64;
constexpr int MAXN = 256;
constexpr int Q_BYTES = 128 * HD * 2;                 // 16 KB
constexpr int KV_BYTES = MAXN * HD * 2;               // 32 KB
constexpr int P_KBLOCK_BYTES = 128 * 128;
constexpr int OFF_Q = 0;
constexpr int OFF_K = Q_BYTES;
constexpr int OFF_P = 0;                              // overlays Q + K
constexpr int OFF_V = 4 * P_KBLOCK_BYTES;             // 64 KB
constexpr int OFF_QAUG = OFF_V + KV_BYTES;            // 128 rows x 32 B, no swizzle
constexpr int OFF_KAUG = OFF_QAUG + 128 * 32;         // 256 rows x 32 B, no swizzle
constexpr int OFF_SCRATCH = OFF_QAUG;                 // full-P output: 4 x 4 KB transpose scratch on top of the (dead) mask operands
constexpr int OFF_CLS = OFF_SCRATCH;                  // CLS row staging [256] floats: first KB of warp 0's scratch (written out before warp 0 transposes)
constexpr int OFF_BAR = OFF_SCRATCH + 4 * 4096;
static_assert(OFF_KAUG + MAXN * 32 <= OFF_BAR, "mask operands fit under the transpose scratch");
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int THREADS = 160;
static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two attention CTAs per SM");
}  // namespace attn2

struct Attn2Params {
    const float* key_bias;   // [B,N] or null
    __nv_bfloat16* out;      // [B,N,H*64]
    float* cls_rows;         // [B,H,N] or null
    float* attn;             // [B,H,N,N] or null
    int B, N, H;
    float scale, scale_log2;
    unsigned long long* trace;   // debug: per-CTA phase timestamps (vtc_debug_set_trace), normally null
    int ts_mode;                 // 1: P stays in TMEM (tcgen05.st) and feeds the second MMA as its TMEM A operand
};

// One 32-column chunk of pass 2: e = 2^(s*sc - m*sc) with packed fp32x2 FMAs, row-sum partials, bf16 pack, swizzled STS.
template <bool CLS>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&cur)[32], int c, int N, uint64_t sc2, uint64_t negm2, uint64_t& sum2,
                                              uint8_t* p_row, int r_local, float* cls_s, bool is_cls_thread, uint32_t t_p, bool ts) {
    uint32_t pk[16];
    const int nvalid = N - c * 32;                    // > 0; >= 32 for all but the last chunk (warp-uniform)
    if (nvalid >= 32) {
        // All 32 exponentials are issued back to back before anything consumes them: MUFU.EX2 retires one warp
        // instruction per 8 clocks per SM sub-partition (measured, tools/ubench/pipes.cu), so the XU pipe is the bound of
        // this loop and must never wait on the FADD2 / F2FP consumers (which then run while the other warp owns the XU).
        float e[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            float x0, x1;
            unpack2(fma2(pack2u(cur[j], cur[j + 1]), sc2, negm2), x0, x1);
            e[j] = ex2_approx(x0);
            e[j + 1] = ex2_approx(x1);
        }
        uint64_t sa = pack2(0.f, 0.f), sb = pack2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            sa = add2(sa, pack2(e[j], e[j + 1]));
            sb = add2(sb, pack2(e[j + 2], e[j + 3]));
            pk[j >> 1] = pack_bf16x2(e[j], e[j + 1]);
            pk[(j >> 1) + 1] = pack_bf16x2(e[j + 2], e[j + 3]);
        }
        sum2 = add2(sum2, add2(sa, sb));
        if (CLS) {
            if (is_cls_thread) {
#pragma unroll
                for (int j = 0; j < 32; ++j) cls_s[c * 32 + j] = e[j];
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            if (j >= nvalid) {                        // padding keys: no exponentials at all
                pk[j >> 1] = 0u;
                continue;
            }
            float x0, x1;
            unpack2(fma2(pack2u(cur[j], cur[j + 1]), sc2, negm2), x0, x1);
            const float e0 = ex2_approx(x0);
            const float e1 = (j + 1 < nvalid) ? ex2_approx(x1) : 0.f;
            sum2 = add2(sum2, pack2(e0, e1));
            if (CLS) {
                if (is_cls_thread) { cls_s[c * 32 + j] = e0; cls_s[c * 32 + j + 1] = e1; }
            }
            pk[j >> 1] = pack_bf16x2(e0, e1);
        }
    }
    if (ts) {                  // P stays in TMEM: bf16 pairs, 16 columns per 32-key chunk, on top of the S columns already consumed
        tmem_st_32x32b_x16(t_p + c * 16, pk);
        return;
    }
    uint8_t* kblk = p_row + (c >> 1) * attn2::P_KBLOCK_BYTES;
    const int g0 = (c & 1) * 4;
#pragma unroll
    for (int g = 0; g < 4; ++g)
        st_u4(kblk + (((g0 + g) ^ (r_local & 7)) * 16), make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]));
}

// Pass 2 over all chunks; TMEM loads run one chunk ahead in two ping-pong register buffers (no copies).
template <bool CLS>
__device__ __forceinline__ void softmax_pass2(uint32_t t_s, int nchunks, int N, float sc, float neg_m, uint8_t* p_row, int r_local,
                                              float* cls_s, bool is_cls_thread, float& sum_out, uint32_t t_p, bool ts) {
    const uint64_t sc2 = pack2(sc, sc), negm2 = pack2(neg_m, neg_m);
    uint64_t sum2 = pack2(0.f, 0.f);
    uint32_t ra[32], rb[32];
    tmem_ld_32x32b_x32(t_s, ra);
    for (int c = 0; c < nchunks; c += 2) {
        tmem_ld_wait();
        if (c + 1 < nchunks) tmem_ld_32x32b_x32(t_s + (c + 1) * 32, rb);
        softmax_chunk<CLS>(ra, c, N, sc2, negm2, sum2, p_row, r_local, cls_s, is_cls_thread, t_p, ts);
        if (c + 1 < nchunks) {
            tmem_ld_wait();
            if (c + 2 < nchunks) tmem_ld_32x32b_x32(t_s + (c + 2) * 32, ra);
            softmax_chunk<CLS>(rb, c + 1, N, sc2, negm2, sum2, p_row, r_local, cls_s, is_cls_thread, t_p, ts);
        }
    }
    float s0, s1;
    unpack2(sum2, s0, s1);
    sum_out = s0 + s1;
}

__device__ __forceinline__ float chunk_max(const uint32_t (&cur)[32], int c, int N, float m) {
    if (c * 32 + 32 <= N) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c * 32 + j < N) m = fmaxf(m, __uint_as_float(cur[j]));
    }
    return m;
}

__global__ void __launch_bounds__(attn2::THREADS, 2)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Attn2Params p) {
    using namespace attn2;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* bar_qk = bars + 0;
    uint64_t* bar_v = bars + 1;
    uint64_t* s_full = bars + 2;
    uint64_t* p_full = bars + 3;
    uint64_t* o_full = bars + 4;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);
    float* cls_s = reinterpret_cast<float*>(smem + OFF_CLS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int N = p.N;
    const int ntiles = (N + 127) >> 7;
    const int mt = blockIdx.x % ntiles;
    const int bh = blockIdx.x / ntiles;
    const int b = bh / p.H;
    const int h = bh - b * p.H;
    const int NP = (N + 15) & ~15;
    const bool has_bias = p.key_bias != nullptr;
    unsigned long long* tr = p.trace ? p.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
    auto stamp = [&](int slot) {
        if (tr) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            tr[slot] = t;
        }
    };
    if (threadIdx.x == 0) stamp(0);

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            mbar_init(bar_qk, 1);
            mbar_init(bar_v, 1);
            mbar_init(s_full, 1);
            mbar_init(p_full, 128);
            mbar_init(o_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 256);
    } else if (has_bias) {
        // augmented K-step operands (no-swizzle K-major core matrices: 8 rows x 16 B, LBO 128 B, SBO 256 B)
        const float* kb = p.key_bias + static_cast<size_t>(b) * N;
        const float inv_scale = 1.0f / p.scale;
        {
            const int r = threadIdx.x;                       // 0..127: query row of this tile
            const int row = mt * 128 + r;
            const float flag = (row < N && kb[row] == 0.f) ? 1.0f : 0.0f;
            uint8_t* dst = smem + OFF_QAUG + (r >> 3) * 256 + (r & 7) * 16;
            st_u4(dst, make_uint4(pack_bf16x2(flag, 0.f), 0u, 0u, 0u));
            st_u4(dst + 128, make_uint4(0u, 0u, 0u, 0u));
        }
        for (int j = threadIdx.x; j < MAXN; j += 128) {
            const float v = (j < N) ? kb[j] * inv_scale : 0.f;
            uint8_t* dst = smem + OFF_KAUG + (j >> 3) * 256 + (j & 7) * 16;
            st_u4(dst, make_uint4(pack_bf16x2(v, 0.f), 0u, 0u, 0u));
            st_u4(dst + 128, make_uint4(0u, 0u, 0u, 0u));
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (threadIdx.x == 0) stamp(1);

    if (warp == 4) {
        if (lane == 0) {
            const int D = p.H * HD;
            mbar_arrive_expect_tx(bar_qk, Q_BYTES + KV_BYTES);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(smem + OFF_Q)), "l"(reinterpret_cast<uint64_t>(&tmQ)), "r"(smem_u32(bar_qk)), "r"(h * HD), "r"(mt * 128), "r"(b)
                : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(smem + OFF_K)), "l"(reinterpret_cast<uint64_t>(&tmKV)), "r"(smem_u32(bar_qk)), "r"(D + h * HD), "r"(0), "r"(b)
                : "memory");
            mbar_arrive_expect_tx(bar_v, KV_BYTES);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(smem + OFF_V)), "l"(reinterpret_cast<uint64_t>(&tmKV)), "r"(smem_u32(bar_v)), "r"(2 * D + h * HD), "r"(0), "r"(b)
                : "memory");

            mbar_wait(bar_qk, 0);
            stamp(6);
            tc_fence_after();
            const uint32_t idesc_s = make_idesc_bf16(128, NP, 0, 0);
            const uint32_t q_addr = smem_u32(smem + OFF_Q);
            const uint32_t k_addr = smem_u32(smem + OFF_K);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
                umma_bf16(tmem_base, make_smem_desc_sw128(q_addr + k * 32, 1024, 16), make_smem_desc_sw128(k_addr + k * 32, 1024, 16),
                          idesc_s, k != 0 ? 1u : 0u);
            if (has_bias)
                umma_bf16(tmem_base, make_smem_desc(smem_u32(smem + OFF_QAUG), 256, 128, 0), make_smem_desc(smem_u32(smem + OFF_KAUG), 256, 128, 0),
                          idesc_s, 1u);
            umma_commit(s_full);

            mbar_wait(bar_v, 0);
            mbar_wait(p_full, 0);
            tc_fence_after();
            const uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
            const uint32_t v_addr = smem_u32(smem + OFF_V);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            const int ksteps = NP / 16;
            if (p.ts_mode) {
                for (int j = 0; j < ksteps; ++j)
                    umma_bf16_ts(tmem_base + 128, tmem_base + 8 * j, make_smem_desc_sw128(v_addr + j * 2048, 1024, 1024), idesc_o, j != 0 ? 1u : 0u);
            } else {
                for (int j = 0; j < ksteps; ++j)
                    umma_bf16(tmem_base, make_smem_desc_sw128(p_addr + (j >> 2) * P_KBLOCK_BYTES + (j & 3) * 32, 1024, 16),
                              make_smem_desc_sw128(v_addr + j * 2048, 1024, 1024), idesc_o, j != 0 ? 1u : 0u);
            }
            umma_commit(o_full);
        }
        __syncwarp();
    } else {
        const int quarter = warp;
        const int r_local = quarter * 32 + lane;
        const int row = mt * 128 + r_local;
        const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const int nchunks = (N + 31) >> 5;
        const float sc = p.scale_log2;
        const bool warp_active = (mt * 128 + quarter * 32) < N;      // warp-uniform: any valid query row in this warp?

        mbar_wait(s_full, 0);
        if (threadIdx.x == 0) stamp(2);
        tc_fence_after();
        float inv = 0.f;
        if (warp_active) {
            // pass 1: row max of the raw accumulator (scale > 0, the mask is already inside S)
            float m = -INFINITY;
            {
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(t_s, ra);
                for (int c = 0; c < nchunks; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < nchunks) tmem_ld_32x32b_x32(t_s + (c + 1) * 32, rb);
                    m = chunk_max(ra, c, N, m);
                    if (c + 1 < nchunks) {
                        tmem_ld_wait();
                        if (c + 2 < nchunks) tmem_ld_32x32b_x32(t_s + (c + 2) * 32, ra);
                        m = chunk_max(rb, c + 1, N, m);
                    }
                }
            }
            const float neg_m = -m * sc;
            uint8_t* p_row = smem + OFF_P + r_local * 128;
            float sum;
            const bool cls_warp = (mt == 0) && (quarter == 0) && (p.cls_rows != nullptr);     // warp-uniform
            const bool ts = p.ts_mode != 0;
            if (cls_warp) softmax_pass2<true>(t_s, nchunks, N, sc, neg_m, p_row, r_local, cls_s, lane == 0, sum, t_s, ts);
            else softmax_pass2<false>(t_s, nchunks, N, sc, neg_m, p_row, r_local, cls_s, false, sum, t_s, ts);
            if (p.ts_mode) tmem_st_wait();
            inv = 1.0f / sum;
            if (cls_warp) {
                const float inv0 = __shfl_sync(0xffffffffu, inv, 0);
                __syncwarp();
                float* dst = p.cls_rows + (static_cast<size_t>(b) * p.H + h) * N;
                for (int j = lane; j < N; j += 32) dst[j] = cls_s[j] * inv0;
                __syncwarp();          // the staging area is the first KB of this warp's transpose scratch
            }
            if (p.attn != nullptr) {
                // pass 3 (on request): normalised fp32 P rows; CTA-uniform branch, only the stores are predicated
                // (the S = Q K^T MMAs have retired, so the mask operands under the scratch are dead)
                float* scratch = reinterpret_cast<float*>(smem + OFF_SCRATCH) + quarter * 1024;
                const int row0 = mt * 128 + quarter * 32;
                float* dst = p.attn + ((static_cast<size_t>(b) * p.H + h) * N + row0) * N;
                for (int c = 0; c < nchunks; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_s + c * 32, r);
                    tmem_ld_wait();
                    float pv[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) pv[j] = ex2_approx(fmaf(__uint_as_float(r[j]), sc, neg_m)) * inv;
                    store_rows_coalesced(scratch, pv, dst + c * 32, static_cast<size_t>(N), N - row0, N - c * 32);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
        }
        if (threadIdx.x == 0) stamp(3);
        mbar_arrive(p_full);

        mbar_wait(o_full, 0);
        if (threadIdx.x == 0) stamp(4);
        tc_fence_after();
        if (warp_active) {
            uint32_t o0[32], o1[32];
            const uint32_t t_o = t_s + (p.ts_mode ? 128u : 0u);
            tmem_ld_32x32b_x32(t_o, o0);
            tmem_ld_32x32b_x32(t_o + 32, o1);
            tmem_ld_wait();
            if (row < N) {
                __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * N + row) * (p.H * HD) + h * HD;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    st_u4(dst + 8 * g, make_uint4(pack_bf16x2(__uint_as_float(o0[8 * g]) * inv, __uint_as_float(o0[8 * g + 1]) * inv),
                                                  pack_bf16x2(__uint_as_float(o0[8 * g + 2]) * inv, __uint_as_float(o0[8 * g + 3]) * inv),
                                                  pack_bf16x2(__uint_as_float(o0[8 * g + 4]) * inv, __uint_as_float(o0[8 * g + 5]) * inv),
                                                  pack_bf16x2(__uint_as_float(o0[8 * g + 6]) * inv, __uint_as_float(o0[8 * g + 7]) * inv)));
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    st_u4(dst + 32 + 8 * g, make_uint4(pack_bf16x2(__uint_as_float(o1[8 * g]) * inv, __uint_as_float(o1[8 * g + 1]) * inv),
                                                       pack_bf16x2(__uint_as_float(o1[8 * g + 2]) * inv, __uint_as_float(o1[8 * g + 3]) * inv),
                                                       pack_bf16x2(__uint_as_float(o1[8 * g + 4]) * inv, __uint_as_float(o1[8 * g + 5]) * inv),
                                                       pack_bf16x2(__uint_as_float(o1[8 * g + 6]) * inv, __uint_as_float(o1[8 * g + 7]) * inv)));
            }
        }
    }
    if (threadIdx.x == 0) stamp(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 256);
    if (threadIdx.x == 0) stamp(7);
}


int attention(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_out, int batch, int n_tokens, int heads,
              float scale, cudaStream_t stream, int reverse, const void* aug) {
    VTC_REQUIRE(qkv && out, VTC_ERR_ARG, "attention: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0, VTC_ERR_SHAPE, "attention: bad shape");
    VTC_REQUIRE(scale > 0.f, VTC_ERR_ARG, "attention: scale must be positive");
    if (attn_out == nullptr)              // fast path: two-group column-split kernel (attention_cs.cu), any sequence length
        return attention_cs(qkv, key_bias, out, cls_rows, batch, n_tokens, heads, scale, stream, reverse, nullptr, aug);
    if (n_tokens > attn2::MAXN)           // full P of long sequences: KV-blocked two-sweep kernel (attention_kv.cu)
        return attention_kv(qkv, key_bias, out, cls_rows, attn_out, batch, n_tokens, heads, scale, false, stream, reverse);
    // full P, N <= 256: the whole row of scores sits in TMEM, the normalised probabilities are a third read of it
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int D = heads * attn2::HD;
    uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)n_tokens, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)3 * D * 2, (uint64_t)n_tokens * 3 * D * 2};
    CUtensorMap tmQ, tmKV;
    uint32_t boxkv[3] = {attn2::HD, attn2::MAXN, 1};
    uint32_t boxq[3] = {attn2::HD, 128, 1};
    rc = make_tmap_bf16(&tmKV, qkv, 3, dims, strides, boxkv);
    if (rc != VTC_OK) return rc;
    rc = make_tmap_bf16(&tmQ, qkv, 3, dims, strides, boxq);
    if (rc != VTC_OK) return rc;
    static SmemOptIn optin;
    if ((rc = optin.ensure(reinterpret_cast<const void*>(attention2_kernel), attn2::SMEM_BYTES, true)) != VTC_OK) return rc;
    Attn2Params p{key_bias, static_cast<__nv_bfloat16*>(out), cls_rows, attn_out, batch, n_tokens, heads, scale, scale * 1.4426950408889634f,
                  g_attn_trace, 0};
    const int ntiles = (n_tokens + 127) / 128;
    attention2_kernel<<<batch * heads * ntiles, attn2::THREADS, attn2::SMEM_BYTES, stream>>>(tmQ, tmKV, p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- head mean of the full P (predict.py:189-190) -------------------------------------------------------
__global__ void head_mean_kernel(const float* __restrict__ attn, float* __restrict__ mean, int H, size_t nn, size_t total) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const float inv = 1.0f / H;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t b = i / nn, e = i - b * nn;
        const float* src = attn + b * H * nn + e;
        float s = 0.f;
        for (int hh = 0; hh < H; ++hh) s += __ldg(src + hh * nn);
        mean[i] = s * inv;
    }
}

int head_mean(const float* attn_in, float* mean, int batch, int heads, int n_tokens, cudaStream_t stream) {
    VTC_REQUIRE(attn_in && mean, VTC_ERR_ARG, "head_mean: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0, VTC_ERR_SHAPE, "head_mean: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const size_t nn = static_cast<size_t>(n_tokens) * n_tokens;
    const size_t total = nn * batch;
    size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    head_mean_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(attn_in, mean, heads, nn, total);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- head mean of the packed P written by attention_cs ---------------------------------------------------------------
//   mean[b,r,k] = 1/H sum_h einv[b,h,r] * 2^(mtab[b,h,r,k/32] - mfin[b,h,r]) * E[b,h,r,k]      (MTAB: several key blocks)
//   mean[b,r,k] = 1/H sum_h einv[b,h,r] * E[b,h,r,k]                                             (one key block)
// One warp per (image, query row); per 256-key segment a lane owns 8 consecutive keys (one 16-byte load per head).  Heads
// are added in order, so the result is bit-reproducible.  A segment leaves through a per-warp staging line because
// [B,N,N] rows (N odd) are not 16-byte aligned: the global stores are 32 consecutive floats per instruction.
constexpr int HMP_WARPS = 8;
// OPERAND: the row leaves as the bf16 rollout operand (ops.h: N values | zero padding | fp32 sum of the rounded values) with
// one 16-byte store per lane; otherwise as fp32 [B,N,N] through the staging line.
template <bool MTAB, bool OPERAND>
__global__ void __launch_bounds__(HMP_WARPS * 32)
head_mean_packed_kernel(const uint4* __restrict__ e, const float* __restrict__ mtab, const float* __restrict__ mfin,
                        const float* __restrict__ einv, float* __restrict__ mean, uint4* __restrict__ operand, int B, int H, int N, int ld, int ldr) {
    __shared__ float stage[HMP_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vec_per_row = ld >> 3;                     // uint4 per row
    const int chunks_per_row = ld >> 5;
    const int nrows = B * N;
    const float invh = 1.0f / static_cast<float>(H);
    float* st = stage[warp];
    for (int rid = blockIdx.x * HMP_WARPS + warp; rid < nrows; rid += gridDim.x * HMP_WARPS) {
        const int b = rid / N, r = rid - b * N;
        const size_t row0 = static_cast<size_t>(b) * H * N + r;          // (b, h = 0, r); + h * N per head
        float* dst = OPERAND ? nullptr : mean + static_cast<size_t>(rid) * N;
        uint4* odst = OPERAND ? operand + static_cast<size_t>(rid) * (ldr >> 3) : nullptr;
        float rsum = 0.f;
        for (int seg = 0; seg * 256 < (OPERAND ? ldr : N); ++seg) {
            const int v = seg * 32 + lane;                // my uint4 of the row; keys 8 v .. 8 v + 7, chunk v / 4
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            if (v < vec_per_row) {
                // four heads per batch: all eight loads are issued before the first fused multiply-add consumes one
                for (int h0 = 0; h0 < H; h0 += 4) {
                    float w[4];
                    uint4 q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool ok = h0 + u < H;
                        const size_t row = row0 + static_cast<size_t>(ok ? h0 + u : H - 1) * N;
                        w[u] = __ldg(einv + row);
                        if (MTAB) w[u] *= exp2f(__ldg(mtab + row * chunks_per_row + (v >> 2)) - __ldg(mfin + row));
                        if (!ok) w[u] = 0.f;
                        q[u] = ld_stream_u4(e + row * vec_per_row + v);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[0] = fmaf(w[u], __uint_as_float(q[u].x << 16), acc[0]);
                        acc[1] = fmaf(w[u], __uint_as_float(q[u].x & 0xffff0000u), acc[1]);
                        acc[2] = fmaf(w[u], __uint_as_float(q[u].y << 16), acc[2]);
                        acc[3] = fmaf(w[u], __uint_as_float(q[u].y & 0xffff0000u), acc[3]);
                        acc[4] = fmaf(w[u], __uint_as_float(q[u].z << 16), acc[4]);
                        acc[5] = fmaf(w[u], __uint_as_float(q[u].z & 0xffff0000u), acc[5]);
                        acc[6] = fmaf(w[u], __uint_as_float(q[u].w << 16), acc[6]);
                        acc[7] = fmaf(w[u], __uint_as_float(q[u].w & 0xffff0000u), acc[7]);
                    }
                }
            }
            if (OPERAND) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = (8 * v + 2 * j < N) ? bf16_round(acc[2 * j] * invh) : 0.f;          // columns >= N: zero padding
                    const float c = (8 * v + 2 * j + 1 < N) ? bf16_round(acc[2 * j + 1] * invh) : 0.f;
                    rsum += a + c;
                    pk[j] = pack_bf16x2(a, c);
                }
                const bool last_seg = (seg + 1) * 256 >= ldr;
                if (last_seg) {                    // the row sum (of the rounded values) rides in the last four bytes of the row
                    rsum = warp_sum(rsum);
                    if (v == (ldr >> 3) - 1) pk[3] = __float_as_uint(rsum);
                }
                if (v < (ldr >> 3)) odst[v] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            } else {
                float4* s4 = reinterpret_cast<float4*>(st + lane * 8);
                s4[0] = make_float4(acc[0] * invh, acc[1] * invh, acc[2] * invh, acc[3] * invh);
                s4[1] = make_float4(acc[4] * invh, acc[5] * invh, acc[6] * invh, acc[7] * invh);
                __syncwarp();
                const int nseg = min(256, N - seg * 256);
                for (int c = lane; c < nseg; c += 32) dst[seg * 256 + c] = st[c];
                __syncwarp();
            }
        }
    }
}

static int head_mean_packed_launch(const PackedP& pk, float* mean, void* operand, int batch, int heads, int n_tokens, int ld, cudaStream_t stream) {
    VTC_REQUIRE(pk.e && pk.einv && (mean || operand) && ((pk.mtab == nullptr) == (pk.mfin == nullptr)), VTC_ERR_ARG, "head_mean_packed: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0 && ld >= n_tokens && ld % 32 == 0, VTC_ERR_SHAPE, "head_mean_packed: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int nrows = batch * n_tokens;
    int blocks = cdiv(nrows, HMP_WARPS);
    const int cap = device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const int ldr = rollout_operand_ld(n_tokens);
    const uint4* e = static_cast<const uint4*>(pk.e);
    uint4* op = static_cast<uint4*>(operand);
    if (operand) VTC_REQUIRE((reinterpret_cast<uintptr_t>(operand) & 15) == 0 && ldr <= ld + 8, VTC_ERR_ARG, "head_mean_packed: operand alignment / stride");
    if (pk.mtab && operand) head_mean_packed_kernel<true, true><<<blocks, HMP_WARPS * 32, 0, stream>>>(e, pk.mtab, pk.mfin, pk.einv, nullptr, op, batch, heads, n_tokens, ld, ldr);
    else if (pk.mtab) head_mean_packed_kernel<true, false><<<blocks, HMP_WARPS * 32, 0, stream>>>(e, pk.mtab, pk.mfin, pk.einv, mean, nullptr, batch, heads, n_tokens, ld, ldr);
    else if (operand) head_mean_packed_kernel<false, true><<<blocks, HMP_WARPS * 32, 0, stream>>>(e, nullptr, nullptr, pk.einv, nullptr, op, batch, heads, n_tokens, ld, ldr);
    else head_mean_packed_kernel<false, false><<<blocks, HMP_WARPS * 32, 0, stream>>>(e, nullptr, nullptr, pk.einv, mean, nullptr, batch, heads, n_tokens, ld, ldr);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}
int head_mean_packed(const PackedP& pk, float* mean, int batch, int heads, int n_tokens, int ld, cudaStream_t stream) {
    return head_mean_packed_launch(pk, mean, nullptr, batch, heads, n_tokens, ld, stream);
}
int head_mean_packed_operand(const PackedP& pk, void* operand, int batch, int heads, int n_tokens, int ld, cudaStream_t stream) {
    return head_mean_packed_launch(pk, nullptr, operand, batch, heads, n_tokens, ld, stream);
}

// scratch layout of attention_mean: E | mtab | mfin | einv, every segment 256-byte aligned
static PackedP carve_packed(void* scratch, int batch, int n_tokens, int heads, size_t* total) {
    const size_t rows = static_cast<size_t>(batch) * heads * n_tokens, ld = attention_packed_ld(n_tokens);
    uint8_t* base = static_cast<uint8_t*>(scratch);
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* q = base ? base + off : nullptr; off += align_up(bytes, 256); return q; };
    PackedP pk;
    const bool multi = n_tokens > kAttentionSingleBlockKeys;      // several key blocks: per-chunk reference maxima (attention_cs.cu)
    pk.e = take(rows * ld * 2);
    pk.einv = reinterpret_cast<float*>(take(rows * 4));
    pk.mtab = multi ? reinterpret_cast<float*>(take(rows * (ld / 32) * 4)) : nullptr;
    pk.mfin = multi ? reinterpret_cast<float*>(take(rows * 4)) : nullptr;
    *total = off;
    return pk;
}
size_t attention_mean_scratch_bytes(int batch, int n_tokens, int heads) {
    if (batch <= 0 || n_tokens <= 0 || heads <= 0) return 0;
    size_t total = 0;
    carve_packed(nullptr, batch, n_tokens, heads, &total);
    return total;
}

int attention_mean(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* scratch, size_t scratch_bytes,
                   int batch, int n_tokens, int heads, float scale, cudaStream_t stream, int reverse) {
    return attention_mean_operand(qkv, key_bias, out, cls_rows, attn_mean, nullptr, scratch, scratch_bytes, batch, n_tokens, heads, scale, stream, reverse, nullptr);
}

int attention_mean_operand(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* operand, void* scratch,
                           size_t scratch_bytes, int batch, int n_tokens, int heads, float scale, cudaStream_t stream, int reverse, const void* aug) {
    VTC_REQUIRE(qkv && out && (attn_mean || operand) && scratch, VTC_ERR_ARG, "attention_mean: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0, VTC_ERR_SHAPE, "attention_mean: bad shape");
    VTC_REQUIRE(n_tokens <= kAttentionFusedMeanMaxTokens, VTC_ERR_SHAPE, "attention_mean: %d tokens > %d", n_tokens, kAttentionFusedMeanMaxTokens);
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, VTC_ERR_WORKSPACE, "attention_mean: scratch must be 256-byte aligned");
    size_t need = 0;
    const PackedP pk = carve_packed(scratch, batch, n_tokens, heads, &need);
    VTC_REQUIRE(scratch_bytes >= need, VTC_ERR_WORKSPACE, "attention_mean: scratch %zu bytes < required %zu", scratch_bytes, need);
    int rc = attention_cs(qkv, key_bias, out, cls_rows, batch, n_tokens, heads, scale, stream, reverse, &pk, aug);
    if (rc != VTC_OK) return rc;
    if (attn_mean && (rc = head_mean_packed(pk, attn_mean, batch, heads, n_tokens, attention_packed_ld(n_tokens), stream)) != VTC_OK) return rc;
    if (operand && (rc = head_mean_packed_operand(pk, operand, batch, heads, n_tokens, attention_packed_ld(n_tokens), stream)) != VTC_OK) return rc;
    return VTC_OK;
}

}  // namespace vtc

extern "C" {
// debug hook (not part of include/vtc.h): device buffer of 8 x uint64 per attention CTA receiving %globaltimer stamps
__attribute__((visibility("default"))) void vtc_debug_set_attention_trace(void* buf) { vtc::g_attn_trace = static_cast<unsigned long long*>(buf); }
int vtc_attention(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int32_t batch, int32_t n_tokens,
                  int32_t heads, float scale, void* stream) {
    return vtc::attention(qkv, key_bias, out, cls_rows, attn, batch, n_tokens, heads, scale, static_cast<cudaStream_t>(stream), 0);
}
int vtc_attention_masked(const void* qkv, const float* key_bias, const void* mask_operands, void* out, float* cls_rows, int32_t batch, int32_t n_tokens,
                         int32_t heads, float scale, void* stream) {
    return vtc::attention(qkv, key_bias, out, cls_rows, nullptr, batch, n_tokens, heads, scale, static_cast<cudaStream_t>(stream), 0, mask_operands);
}
size_t vtc_attention_mean_scratch_bytes(int32_t batch, int32_t n_tokens, int32_t heads) {
    return vtc::attention_mean_scratch_bytes(batch, n_tokens, heads);
}
int vtc_attention_mean(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* scratch, size_t scratch_bytes,
                       int32_t batch, int32_t n_tokens, int32_t heads, float scale, void* stream) {
    return vtc::attention_mean(qkv, key_bias, out, cls_rows, attn_mean, scratch, scratch_bytes, batch, n_tokens, heads, scale,
                               static_cast<cudaStream_t>(stream), 0);
}
int vtc_attention_mean_operand(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* operand, void* scratch,
                               size_t scratch_bytes, int32_t batch, int32_t n_tokens, int32_t heads, float scale, void* stream) {
    return vtc::attention_mean_operand(qkv, key_bias, out, cls_rows, attn_mean, operand, scratch, scratch_bytes, batch, n_tokens, heads, scale,
                                       static_cast<cudaStream_t>(stream), 0);
}
int vtc_head_mean(const float* attn, float* mean, int32_t batch, int32_t heads, int32_t n_tokens, void* stream) {
    return vtc::head_mean(attn, mean, batch, heads, n_tokens, static_cast<cudaStream_t>(stream));
}
}
