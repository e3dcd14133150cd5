// KV-blocked fused attention for any sequence length (vit_model.py:113-137 at N = 577 / 785 tokens: BASELINE configs 4
// and 5) and for the split-bf16 "fp32 mode".  Same math and outputs as attention.cu: O = softmax(Q K^T * scale + mask) V,
// the CLS query row P[b,h,0,:] and, on request, the whole normalised P.
//
// One persistent CTA per SM holds TWO independent pipelines ("groups"); a group owns 256 TMEM columns, its own Q / K / V
// shared-memory tiles and three warp roles:
//   4 softmax warps   one query row per thread (TMEM lane = row of the 128-row query tile)
//   1 producer warp   TMA loads of Q (once per item) and of the K / V blocks (KB <= 192 keys each)
//   1 MMA warp        S_j = Q K_j^T  ->  [softmax]  ->  O += P_j V_j, block after block, one elected thread
// A work item is (image, head, 128-query tile).  The two groups run different items, so while one group waits for its
// MMAs the other one keeps the MUFU / TMEM ports busy.
//
// Softmax is single pass with a lazily raised running maximum (as in attention3_kernel): S_j is read from TMEM once,
// exponentials are taken against the running maximum m, which is only raised when a chunk exceeds it by more than 2^8;
// in that (rare) event the bf16 P chunks already stored, the row sum, the staged CLS row and the O accumulator in TMEM
// are rescaled by 2^(m_old - m_new), an exact power of two (m moves by whole octaves).  P never touches shared memory: it overwrites the consumed S columns as bf16 pairs
// (tcgen05.st) and feeds the second MMA as its TMEM A operand.
//
// The reference mask -100*min(v_i + v_j, 1) (vit_model.py:348-361) is applied in registers: key j gets key_bias[j] on
// query rows with v_i = 0; rows with v_i = 1 stay unmasked (their uniform -100 is softmax-invariant).
//
// Full P (attn != NULL): a second sweep recomputes every S_j with the final (m, 1/sum) of the row and writes the
// normalised fp32 probabilities; nothing is kept between the sweeps but two scalars per row.
//
// SPLIT = true (fp32 mode): q, k, v and P are carried as (hi, lo) bf16 pairs, every product a.b is evaluated as
// a_hi.b_hi + a_lo.b_hi + a_hi.b_lo with fp32 accumulation (3 MMAs, relative error ~2^-17 instead of 2^-9).
#include "common.cuh"
#include "ops.h"
#include "tma_host.h"

namespace vtc {

namespace akv {
constexpr int HD = 64;
constexpr int NMAX = 2048;                 // longest sequence (tokens) the staging buffers are sized for
constexpr int THREADS = 384;               // 8 softmax warps, 2 producer warps, 2 MMA warps
constexpr int O_COL = 192;                 // O accumulator: TMEM columns 192..255 of the group's region
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 domain

template <bool SPLIT>
struct Cfg {
    static constexpr int KBMAX = SPLIT ? 128 : 192;
    static constexpr int PARTS = SPLIT ? 2 : 1;
    static constexpr int Q_PART = 128 * 128;                    // 128 rows x 64 bf16
    static constexpr int KV_PART = KBMAX * 128;
    static constexpr int OFF_K = Q_PART * PARTS;
    static constexpr int OFF_V = OFF_K + KV_PART * PARTS;
    static constexpr int GROUP_BYTES = OFF_V + KV_PART * PARTS;
    static constexpr int OFF_CLS = 2 * GROUP_BYTES;             // [2][NMAX] floats: staged CLS row
    static constexpr int OFF_KB = OFF_CLS + 2 * NMAX * 4;       // [2][NMAX] floats: key bias, log2 domain
    static constexpr int OFF_BAR = OFF_KB + 2 * NMAX * 4;
    static constexpr int OFF_SCRATCH = OFF_BAR + 256;           // full-P output: 8 x 4 KB transpose scratch (not in split mode: no room)
    static constexpr int SMEM_BYTES = OFF_SCRATCH + (SPLIT ? 0 : 8 * 4096);
    static constexpr int PLO_COL = 128;                         // SPLIT: low halves of P
    static_assert(SMEM_BYTES <= 232448, "attention_kv smem budget");
    static_assert(GROUP_BYTES % 1024 == 0 && OFF_K % 1024 == 0 && OFF_V % 1024 == 0, "swizzle atoms need 1024-byte tiles");
};

struct Params {
    const float* key_bias;   // [B,N] or null
    __nv_bfloat16* out;      // [B*N, out_stride]; SPLIT: hi at column h*64, lo at column lo_off + h*64
    float* cls_rows;         // [B,H,N] or null
    float* attn;             // [B,H,N,N] or null
    int B, N, H;
    int KB, nb;              // keys per block (multiple of 32), number of blocks
    int out_stride, lo_off;
    float scale_log2;
    int reverse;             // walk the items from the last to the first
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// multiply the bf16 pairs of TMEM columns [col, col+16) of this warp's lanes by f
__device__ __forceinline__ void rescale_p16(uint32_t taddr, float f) {
    uint32_t q[16];
    tmem_ld_32x32b_x16(taddr, q);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float lo = __uint_as_float(q[j] << 16) * f, hi = __uint_as_float(q[j] & 0xffff0000u) * f;
        q[j] = pack_bf16x2(lo, hi);
    }
    tmem_st_32x32b_x16(taddr, q);
}
__device__ __forceinline__ void rescale_f16(uint32_t taddr, float f) {
    uint32_t q[16];
    tmem_ld_32x32b_x16(taddr, q);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) q[j] = __float_as_uint(__uint_as_float(q[j]) * f);
    tmem_st_32x32b_x16(taddr, q);
}

// State of one query row (one thread) across the key blocks of an item.
struct RowState {
    float m;         // running maximum, log2 domain (logit * scale * log2 e + bias)
    uint64_t sum2;   // two partial row sums
};

// One 32-key chunk of sweep 1.  cur = raw S accumulator values (fp32 bits) of this thread's row.
//   c_in_blk: chunk index inside the key block; col0: global key index of the chunk's first column
//   t_s: TMEM address of the group's region for this warp's lanes; have_o: the O accumulator already holds earlier blocks
template <bool SPLIT, bool BIAS>
__device__ __forceinline__ void chunk_sweep1(uint32_t (&cur)[32], int c_in_blk, int col0, int N, float sc, float rb, const float* kb_s,
                                             RowState& st, bool first, uint32_t t_s, bool have_o, float* cls_s, bool cls_thread) {
    const int nvalid = N - col0;                      // > 0, warp-uniform
    const uint64_t sc2 = pack2(sc, sc);
    // ---- logits in the log2 domain (BIAS) / raw accumulator (no bias), and their maximum over the valid keys
    float mc = -INFINITY;
    if (BIAS) {
        const uint64_t rb2 = pack2(rb, rb);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(kb_s + col0 + j);
            float x0, x1, x2, x3;
            unpack2(fma2(pack2u(cur[j], cur[j + 1]), sc2, mul2(pack2(k4.x, k4.y), rb2)), x0, x1);
            unpack2(fma2(pack2u(cur[j + 2], cur[j + 3]), sc2, mul2(pack2(k4.z, k4.w), rb2)), x2, x3);
            cur[j] = __float_as_uint(x0); cur[j + 1] = __float_as_uint(x1);
            cur[j + 2] = __float_as_uint(x2); cur[j + 3] = __float_as_uint(x3);
        }
    }
    if (nvalid >= 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) mc = fmaxf(mc, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) mc = fmaxf(mc, __uint_as_float(cur[j]));
    }
    if (!BIAS) mc *= sc;                              // scale > 0
    // ---- lazily raise the running maximum
    if (first) {
        st.m = mc;
    } else {
        const bool need = (mc - st.m) > RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
            // raise m by a whole number of octaves: the factor is an exact power of two, so rescaling the stored bf16 (hi, lo)
            // pairs, the fp32 O accumulator and the row sum introduces no rounding at all
            const int d = need ? static_cast<int>(ceilf(mc - st.m)) : 0;
            const float f = d >= 127 ? 0.f : __int_as_float((127 - d) << 23);
            st.m += static_cast<float>(d);
            st.sum2 = mul2(st.sum2, pack2(f, f));
            tmem_st_wait();
            for (int cc = 0; cc < c_in_blk; ++cc) {
                rescale_p16(t_s + cc * 16, f);
                if (SPLIT) rescale_p16(t_s + Cfg<SPLIT>::PLO_COL + cc * 16, f);
            }
            if (have_o) {
#pragma unroll
                for (int q = 0; q < 4; ++q) rescale_f16(t_s + O_COL + q * 16, f);
            }
            tmem_st_wait();
            if (cls_thread)
                for (int j = 0; j < col0; ++j) cls_s[j] *= f;
        }
    }
    const float negm = -st.m;
    const uint64_t negm2 = pack2(negm, negm);
    // ---- exponentials, row sum, bf16 P
    uint32_t pk[16];
    uint32_t pl[SPLIT ? 16 : 1];
    if (nvalid >= 32) {
        float e[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            float x0, x1;
            if (BIAS) unpack2(add2(pack2u(cur[j], cur[j + 1]), negm2), x0, x1);
            else unpack2(fma2(pack2u(cur[j], cur[j + 1]), sc2, negm2), x0, x1);
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
            if (SPLIT) {
                pl[j >> 1] = pack_bf16x2(e[j] - __uint_as_float(pk[j >> 1] << 16), e[j + 1] - __uint_as_float(pk[j >> 1] & 0xffff0000u));
                pl[(j >> 1) + 1] = pack_bf16x2(e[j + 2] - __uint_as_float(pk[(j >> 1) + 1] << 16),
                                               e[j + 3] - __uint_as_float(pk[(j >> 1) + 1] & 0xffff0000u));
            }
        }
        st.sum2 = add2(st.sum2, add2(sa, sb));
        if (cls_thread) {
#pragma unroll
            for (int j = 0; j < 32; ++j) cls_s[col0 + j] = e[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            if (j >= nvalid) {                        // padding keys: exact zeros, no exponentials
                pk[j >> 1] = 0u;
                if (SPLIT) pl[j >> 1] = 0u;
                continue;
            }
            float x0, x1;
            if (BIAS) unpack2(add2(pack2u(cur[j], cur[j + 1]), negm2), x0, x1);
            else unpack2(fma2(pack2u(cur[j], cur[j + 1]), sc2, negm2), x0, x1);
            const float e0 = ex2_approx(x0);
            const float e1 = (j + 1 < nvalid) ? ex2_approx(x1) : 0.f;
            st.sum2 = add2(st.sum2, pack2(e0, e1));
            if (cls_thread) {
                cls_s[col0 + j] = e0;
                if (j + 1 < nvalid) cls_s[col0 + j + 1] = e1;
            }
            pk[j >> 1] = pack_bf16x2(e0, e1);
            if (SPLIT) pl[j >> 1] = pack_bf16x2(e0 - __uint_as_float(pk[j >> 1] << 16), e1 - __uint_as_float(pk[j >> 1] & 0xffff0000u));
        }
    }
    tmem_st_32x32b_x16(t_s + c_in_blk * 16, pk);
    if constexpr (SPLIT) tmem_st_32x32b_x16(t_s + Cfg<SPLIT>::PLO_COL + c_in_blk * 16, pl);
}

// One 32-key chunk of sweep 2: normalised probabilities of this thread's row -> global fp32.
// scratch != null: the 32 x 32 block of this warp goes out through the coalescing transpose (dst_blk -> element (warp row 0,
// col0), rows_left = valid rows from there); scratch == null (split mode: no shared memory left): row-wise scalar stores.
template <bool BIAS>
__device__ __forceinline__ void chunk_sweep2(const uint32_t (&cur)[32], int col0, int N, float sc, float rb, const float* kb_s, float negm,
                                             float inv, float* dst_row, bool wr, float* scratch, float* dst_blk, int rows_left) {
    const int nvalid = N - col0;
    float pv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(cur[j]) * sc;
        if (BIAS) x = fmaf(kb_s[(j < nvalid) ? col0 + j : col0], rb, x);
        pv[j] = ex2_approx(x + negm) * inv;
    }
    if (scratch != nullptr) {
        store_rows_coalesced(scratch, pv, dst_blk + col0, static_cast<size_t>(N), rows_left, nvalid);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (wr && j < nvalid) dst_row[col0 + j] = pv[j];
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
attention_kv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Params p) {
    using C = Cfg<SPLIT>;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // role and group of this warp
    const int g = (warp < 8) ? (warp >> 2) : ((warp - 8) & 1);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR) + g * 10;
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;
    uint64_t* k_empty = bars + 3;
    uint64_t* v_full = bars + 4;
    uint64_t* v_empty = bars + 5;
    uint64_t* s_full = bars + 6;     // S_j ready in TMEM
    uint64_t* p_full = bars + 7;     // P_j stored / S_j consumed (128 threads)
    uint64_t* o_full = bars + 8;     // O complete
    uint64_t* o_empty = bars + 9;    // O read out (128 threads)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 20 * 8);
    uint8_t* gsm = smem + g * C::GROUP_BYTES;
    float* cls_s = reinterpret_cast<float*>(smem + C::OFF_CLS) + g * NMAX;
    float* kb_s = reinterpret_cast<float*>(smem + C::OFF_KB) + g * NMAX;

    const int N = p.N, H = p.H, KB = p.KB, nb = p.nb;
    const int qtiles = (N + 127) >> 7;
    const int n_items = p.B * H * qtiles;
    const int first_item = blockIdx.x * 2 + g;
    const int item_stride = gridDim.x * 2;
    const bool want_p = p.attn != nullptr;
    const int D = H * HD;
    const uint32_t kv_bytes = static_cast<uint32_t>(KB) * 128u * C::PARTS;

    if (warp == 10) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            uint64_t* all = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
            for (int gg = 0; gg < 2; ++gg) {
                for (int i = 0; i < 10; ++i) mbar_init(all + gg * 10 + i, (i == 7 || i == 9) ? 128 : 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr + g * 256;

    if (warp == 8 || warp == 9) {
        // ---------------- producer ----------------
        if (lane == 0) {
            uint32_t kidx = 0, vidx = 0, qidx = 0;
            for (int it_ = first_item; it_ < n_items; it_ += item_stride, ++qidx) {
                const int it = p.reverse ? n_items - 1 - it_ : it_;
                const int qt = it % qtiles;
                const int bh = it / qtiles;
                const int b = bh / H, h = bh - b * H;
                mbar_wait(q_empty, (qidx & 1) ^ 1);
                mbar_arrive_expect_tx(q_full, C::Q_PART * C::PARTS);
                tma_load_3d(gsm, &tmQ, q_full, h * HD, qt * 128, b);
                if (SPLIT) tma_load_3d(gsm + C::Q_PART, &tmQ, q_full, 3 * D + h * HD, qt * 128, b);
                for (int pass = 0; pass < (want_p ? 2 : 1); ++pass) {
                    for (int j = 0; j < nb; ++j) {
                        mbar_wait(k_empty, (kidx & 1) ^ 1);
                        mbar_arrive_expect_tx(k_full, kv_bytes);
                        tma_load_3d(gsm + C::OFF_K, &tmKV, k_full, D + h * HD, j * KB, b);
                        if (SPLIT) tma_load_3d(gsm + C::OFF_K + C::KV_PART, &tmKV, k_full, 4 * D + h * HD, j * KB, b);
                        ++kidx;
                        if (pass == 0) {
                            mbar_wait(v_empty, (vidx & 1) ^ 1);
                            mbar_arrive_expect_tx(v_full, kv_bytes);
                            tma_load_3d(gsm + C::OFF_V, &tmKV, v_full, 2 * D + h * HD, j * KB, b);
                            if (SPLIT) tma_load_3d(gsm + C::OFF_V + C::KV_PART, &tmKV, v_full, 5 * D + h * HD, j * KB, b);
                            ++vidx;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 10) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(gsm);
            const uint32_t k_addr = smem_u32(gsm + C::OFF_K);
            const uint32_t v_addr = smem_u32(gsm + C::OFF_V);
            const uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
            uint32_t kidx = 0, vidx = 0, qidx = 0, step = 0;
            bool pending_p = false;      // the S region is still being read by the softmax warps (sweep 2)
            for (int it = first_item; it < n_items; it += item_stride, ++qidx) {
                mbar_wait(q_full, qidx & 1);
                for (int pass = 0; pass < (want_p ? 2 : 1); ++pass) {
                    for (int j = 0; j < nb; ++j) {
                        const int vj = min(KB, N - j * KB);
                        const int nmma = (vj + 15) & ~15;
                        if (pending_p) {
                            mbar_wait(p_full, (step - 1) & 1);
                            pending_p = false;
                        }
                        mbar_wait(k_full, kidx & 1);
                        tc_fence_after();
                        const uint32_t idesc_s = make_idesc_bf16(128, nmma, 0, 0);
#pragma unroll
                        for (int part = 0; part < (SPLIT ? 3 : 1); ++part) {
                            const uint32_t qa = q_addr + ((part == 1) ? C::Q_PART : 0);       // hi, lo, hi
                            const uint32_t ka = k_addr + ((part == 2) ? C::KV_PART : 0);      // hi, hi, lo
#pragma unroll
                            for (int k = 0; k < HD / 16; ++k)
                                umma_bf16(tmem_base, make_smem_desc_sw128(qa + k * 32, 1024, 16), make_smem_desc_sw128(ka + k * 32, 1024, 16),
                                          idesc_s, (part | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(s_full);
                        umma_commit(k_empty);
                        ++kidx;
                        if (pass == (want_p ? 1 : 0) && j == nb - 1) umma_commit(q_empty);
                        if (pass == 0) {
                            mbar_wait(p_full, step & 1);
                            if (j == 0) mbar_wait(o_empty, (qidx & 1) ^ 1);
                            mbar_wait(v_full, vidx & 1);
                            tc_fence_after();
                            const int ksteps = nmma >> 4;
#pragma unroll
                            for (int part = 0; part < (SPLIT ? 3 : 1); ++part) {
                                const uint32_t pa = tmem_base + ((part == 1) ? C::PLO_COL : 0);       // hi, lo, hi
                                const uint32_t va = v_addr + ((part == 2) ? C::KV_PART : 0);          // hi, hi, lo
                                for (int ks = 0; ks < ksteps; ++ks)
                                    umma_bf16_ts(tmem_base + O_COL, pa + 8 * ks, make_smem_desc_sw128(va + ks * 2048, 1024, 1024), idesc_o,
                                                 (j | part | ks) != 0 ? 1u : 0u);
                            }
                            umma_commit(v_empty);
                            ++vidx;
                            if (j == nb - 1) umma_commit(o_full);
                        } else {
                            pending_p = true;
                        }
                        ++step;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------- softmax warps ----------------
        const int quarter = warp & 3;
        const int r_local = quarter * 32 + lane;
        const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float sc = p.scale_log2;
        const bool has_bias = p.key_bias != nullptr;
        uint32_t qidx = 0, step = 0;
        int b_staged = -1;
        for (int it_ = first_item; it_ < n_items; it_ += item_stride, ++qidx) {
            const int it = p.reverse ? n_items - 1 - it_ : it_;
            const int qt = it % qtiles;
            const int bh = it / qtiles;
            const int b = bh / H, h = bh - b * H;
            const int row = qt * 128 + r_local;
            const bool warp_active = (qt * 128 + quarter * 32) < N;
            const bool cls_warp = (qt == 0) && (quarter == 0) && (p.cls_rows != nullptr);
            const bool cls_thread = cls_warp && lane == 0;
            float rb = 1.f;
            if (has_bias) {
                if (b != b_staged) {                  // group-uniform
                    named_bar_sync(1 + g, 128);       // everyone is done with the previous image's bias
                    const float* kb = p.key_bias + static_cast<size_t>(b) * N;
                    const int npad = (N + 31) & ~31;
                    for (int j = threadIdx.x & 127; j < npad; j += 128) kb_s[j] = (j < N) ? kb[j] * 1.4426950408889634f : 0.f;
                    named_bar_sync(1 + g, 128);
                    b_staged = b;
                }
                // a query row that is itself background (v_i = 1) is not masked (vit_model.py:348-361)
                rb = (row < N && kb_s[row] != 0.f) ? 0.f : 1.f;
            }
            RowState st;
            st.m = 0.f;
            st.sum2 = pack2(0.f, 0.f);
            // ---- sweep 1: P_j, row sum, O
            for (int j = 0; j < nb; ++j, ++step) {
                const int vj = min(KB, N - j * KB);
                const int nch = (vj + 31) >> 5;
                mbar_wait(s_full, step & 1);
                tc_fence_after();
                if (warp_active) {
                    uint32_t ra[32], rbuf[32];
                    tmem_ld_32x32b_x32(t_s, ra);
                    for (int c = 0; c < nch; c += 2) {
                        tmem_ld_wait();
                        if (c + 1 < nch) tmem_ld_32x32b_x32(t_s + (c + 1) * 32, rbuf);
                        if (has_bias) chunk_sweep1<SPLIT, true>(ra, c, j * KB + c * 32, N, sc, rb, kb_s, st, (j | c) == 0, t_s, j > 0, cls_s, cls_thread);
                        else chunk_sweep1<SPLIT, false>(ra, c, j * KB + c * 32, N, sc, rb, kb_s, st, (j | c) == 0, t_s, j > 0, cls_s, cls_thread);
                        if (c + 1 < nch) {
                            tmem_ld_wait();
                            if (c + 2 < nch) tmem_ld_32x32b_x32(t_s + (c + 2) * 32, ra);
                            if (has_bias) chunk_sweep1<SPLIT, true>(rbuf, c + 1, j * KB + (c + 1) * 32, N, sc, rb, kb_s, st, false, t_s, j > 0, cls_s, cls_thread);
                            else chunk_sweep1<SPLIT, false>(rbuf, c + 1, j * KB + (c + 1) * 32, N, sc, rb, kb_s, st, false, t_s, j > 0, cls_s, cls_thread);
                        }
                    }
                    tmem_st_wait();
                }
                tc_fence_before();
                mbar_arrive(p_full);
            }
            float s0, s1;
            unpack2(st.sum2, s0, s1);
            const float inv = warp_active ? 1.0f / (s0 + s1) : 0.f;
            if (cls_warp) {
                const float inv0 = __shfl_sync(0xffffffffu, inv, 0);
                __syncwarp();
                float* dst = p.cls_rows + (static_cast<size_t>(b) * H + h) * N;
                for (int j = lane; j < N; j += 32) dst[j] = cls_s[j] * inv0;
                __syncwarp();
            }
            // ---- epilogue: O / rowsum
            mbar_wait(o_full, qidx & 1);
            tc_fence_after();
            if (warp_active) {
                uint32_t o0[32], o1[32];
                tmem_ld_32x32b_x32(t_s + O_COL, o0);
                tmem_ld_32x32b_x32(t_s + O_COL + 32, o1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_empty);
                if (row < N) {
                    __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * N + row) * p.out_stride + h * HD;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t(&o)[32] = half ? o1 : o0;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float a = __uint_as_float(o[8 * q4 + 2 * e]) * inv, c = __uint_as_float(o[8 * q4 + 2 * e + 1]) * inv;
                                hi[e] = pack_bf16x2(a, c);
                                if (SPLIT) lo[e] = pack_bf16x2(a - __uint_as_float(hi[e] << 16), c - __uint_as_float(hi[e] & 0xffff0000u));
                            }
                            st_u4(dst + half * 32 + 8 * q4, make_uint4(hi[0], hi[1], hi[2], hi[3]));
                            if (SPLIT) st_u4(dst + p.lo_off + half * 32 + 8 * q4, make_uint4(lo[0], lo[1], lo[2], lo[3]));
                        }
                    }
                }
            } else {
                tc_fence_before();
                mbar_arrive(o_empty);
            }
            // ---- sweep 2 (on request): normalised P rows
            if (want_p) {
                const float negm = -st.m;
                const bool wr = row < N;
                float* dst_row = p.attn + ((static_cast<size_t>(b) * H + h) * N + (wr ? row : 0)) * N;
                const int row0 = qt * 128 + quarter * 32;
                float* scratch = SPLIT ? nullptr : reinterpret_cast<float*>(smem + C::OFF_SCRATCH) + warp * 1024;
                float* dst_blk = p.attn + ((static_cast<size_t>(b) * H + h) * N + row0) * N;
                for (int j = 0; j < nb; ++j, ++step) {
                    const int vj = min(KB, N - j * KB);
                    const int nch = (vj + 31) >> 5;
                    mbar_wait(s_full, step & 1);
                    tc_fence_after();
                    if (warp_active) {
                        for (int c = 0; c < nch; ++c) {
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(t_s + c * 32, r);
                            tmem_ld_wait();
                            if (has_bias) chunk_sweep2<true>(r, j * KB + c * 32, N, sc, rb, kb_s, negm, inv, dst_row, wr, scratch, dst_blk, N - row0);
                            else chunk_sweep2<false>(r, j * KB + c * 32, N, sc, rb, kb_s, negm, inv, dst_row, wr, scratch, dst_blk, N - row0);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(p_full);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 10) tmem_dealloc(*tmem_ptr, 512);
}

template <bool SPLIT>
static int launch(const void* qkv, const Params& p, cudaStream_t stream) {
    using C = Cfg<SPLIT>;
    const int D = p.H * HD;
    const uint64_t cols = static_cast<uint64_t>(SPLIT ? 6 : 3) * D;
    uint64_t dims[3] = {cols, (uint64_t)p.N, (uint64_t)p.B};
    uint64_t strides[2] = {cols * 2, (uint64_t)p.N * cols * 2};
    CUtensorMap tmQ, tmKV;
    uint32_t boxq[3] = {HD, 128, 1};
    uint32_t boxkv[3] = {HD, static_cast<uint32_t>(p.KB), 1};
    int rc = make_tmap_bf16(&tmQ, qkv, 3, dims, strides, boxq);
    if (rc != VTC_OK) return rc;
    rc = make_tmap_bf16(&tmKV, qkv, 3, dims, strides, boxkv);
    if (rc != VTC_OK) return rc;
    static SmemOptIn optin;
    if ((rc = optin.ensure(reinterpret_cast<const void*>(attention_kv_kernel<SPLIT>), C::SMEM_BYTES)) != VTC_OK) return rc;
    const int items = p.B * p.H * cdiv(p.N, 128);
    int grid = cdiv(items, 2);
    if (grid > device_sm_count()) grid = device_sm_count();
    attention_kv_kernel<SPLIT><<<grid, THREADS, C::SMEM_BYTES, stream>>>(tmQ, tmKV, p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}
}  // namespace akv

int attention_kv(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_out, int batch, int n_tokens, int heads,
                 float scale, bool split, cudaStream_t stream, int reverse) {
    VTC_REQUIRE(qkv && out, VTC_ERR_ARG, "attention: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0, VTC_ERR_SHAPE, "attention: bad shape");
    VTC_REQUIRE(scale > 0.f, VTC_ERR_ARG, "attention: scale must be positive");
    VTC_REQUIRE(n_tokens <= akv::NMAX, VTC_ERR_SHAPE, "attention: %d tokens > %d", n_tokens, akv::NMAX);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int kbmax = split ? akv::Cfg<true>::KBMAX : akv::Cfg<false>::KBMAX;
    akv::Params p{};
    p.key_bias = key_bias;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.cls_rows = cls_rows;
    p.attn = attn_out;
    p.B = batch;
    p.N = n_tokens;
    p.H = heads;
    p.nb = cdiv(n_tokens, kbmax);
    p.KB = (cdiv(n_tokens, p.nb) + 31) & ~31;
    p.out_stride = heads * akv::HD * (split ? 2 : 1);
    p.lo_off = heads * akv::HD;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.reverse = reverse;
    return split ? akv::launch<true>(qkv, p, stream) : akv::launch<false>(qkv, p, stream);
}

}  // namespace vtc

extern "C" int vtc_attention_kv(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int32_t batch,
                                int32_t n_tokens, int32_t heads, float scale, int32_t split, void* stream) {
    return vtc::attention_kv(qkv, key_bias, out, cls_rows, attn, batch, n_tokens, heads, scale, split != 0, static_cast<cudaStream_t>(stream), 0);
}
