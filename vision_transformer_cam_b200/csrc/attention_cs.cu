// Fused attention, fast path (vit_model.py:113-137): O = softmax(Q K^T * scale + mask) V and the CLS query row
// P[b,h,0,:], any sequence length up to 2048 tokens, bf16 operands.  (Full-P output and the split-bf16 fp32 mode
// are served by attention_kv.cu.)
//
// Every score has to come out of TMEM once as fp32 (54 B/clk per SM measured, tools/ubench/tmem.cu), and tcgen05.ld stalls
// the issuing warp for the whole transfer (303 clk per 32x32 block), so the MUFU / FMA work of a warp cannot overlap its own
// loads -- only other warps on the same sub-partition can fill that time.  (Round-2 ablation, DESIGN.md 3.1: with four
// softmax warps per sub-partition the loads ARE hidden -- removing them changes nothing; what is exposed is the softmax
// arithmetic, MUFU-bound while the two groups' softmax phases overlap, on top of a per-item chain of MMA round trips.)
// Hence the shape:
//   * one persistent CTA per SM runs TWO independent pipelines ("groups"), each on its own stream of work items
//     (image, head, 128-query tile), with its own 256 TMEM columns, Q / K / V tiles, producer warp and MMA warp;
//   * inside a group the 128 x KB score block is split by COLUMNS between two sets of four softmax warps (PARTS = 2):
//     warp w works on column range (w / 4) % 2 for the 32 rows of TMEM lane quarter w % 4.
// That puts four softmax warps on every sub-partition, two per group; while one group waits for a tensor-core round trip
// (QK^T of its next block, or P V before the epilogue) the other group's warps own the port.
//   warps 0-7 / 8-15   softmax + epilogue of group 0 / 1
//   warps 16, 17       producers: TMA loads of Q (per item) and of the K / V blocks (<= 208 keys), one stage each -- the
//                      next block is fetched while the current one is in the softmax; they also write the mask operands
//   warps 18, 19       MMA issuers (one elected thread each): QK(s) -> [softmax] -> PV(s) -> QK(s+1) ...
// TMEM region of a group (256 columns): S at 0..207, the O accumulator at 192..255 (it is written by PV only after the
// softmax has consumed S).  The bf16 P of a column half is stored 16 columns into that half's own S range, so S columns
// 0..15 of a block stay intact during the softmax: both halves read the scores of the first 8 keys from there and derive
// the same initial row maximum without talking to each other.
//
// Softmax: single pass over S, exponentials taken against a running maximum m that starts at ceil(max of the first 8
// keys) and is raised, by whole octaves only, when a chunk exceeds it by more than 2^8 -- the stored bf16 P chunks, the
// row sum and (between key blocks) the O accumulator are then rescaled by an exact power of two.  The two warps that
// share a row exchange (m, row sum) through shared memory once per block (one 64-thread named barrier).  P overwrites
// the consumed S columns of the SAME warp as bf16 pairs (tcgen05.st) and is the TMEM A operand of the second MMA; V is
// consumed MN-major exactly as TMA wrote it.
//
// Mask: the reference's -100*min(v_i + v_j, 1) (vit_model.py:348-361) is applied BY THE TENSOR CORE through one extra
// K-step with Q_aug[i] = [v_i == 0] and K_aug[j] = key_bias[j] / scale, so the softmax code has no mask handling and
// background query rows stay unmasked exactly like the reference (their uniform -100 is softmax-invariant).
#include "common.cuh"
#include "ops.h"
#include "tma_host.h"

namespace vtc {

namespace acs {
// Ablation switches for timing experiments (tools/ab_attn_lib.py on libraries built with -DVTC_ACS_ABLATE=k; results are WRONG
// by construction): bit 0 = no softmax arithmetic (a chunk is stored back as loaded), bit 1 = no tcgen05.ld of the scores,
// bit 2 = one P V k-step instead of all, bit 3 = no tcgen05.ld of O in the epilogue, bit 4 = no MUFU.EX2 in full chunks (e = x),
// bit 5 = no maximum pass over full chunks.  0 in every shipped build.
#ifndef VTC_ACS_ABLATE
#define VTC_ACS_ABLATE 0
#endif
constexpr int ABLATE = VTC_ACS_ABLATE;
// The %clock64 timeline (tools/attn_trace_cs.py) is compiled in only on request (tools/build_ablate.sh, digit 7): the stamps' address
// arithmetic costs registers in kernels that sit at their 96-register cap.
#ifndef VTC_ACS_TRACE
#define VTC_ACS_TRACE 0
#endif
constexpr bool TRACE_ALL = VTC_ACS_TRACE != 0;
constexpr int HD = 64;
constexpr int NMAX = 2048;
constexpr int KBMAX = 208;                 // keys per block when the whole sequence fits one block
constexpr int KBLONG = 192;                // keys per block otherwise (S must not reach into the O accumulator)
constexpr int GROUPS = 2;
constexpr int PARTS = 2;                   // column ranges per score block
constexpr int GROUP_WARPS = 4 * PARTS;     // softmax warps of a group
constexpr int SM_WARPS = GROUPS * GROUP_WARPS;
constexpr int THREADS = (SM_WARPS + 2 * GROUPS) * 32;
constexpr int OCOLS = 64 / PARTS;          // O columns each part normalises and stores
constexpr int REGION_COLS = 256;
constexpr int O_COL = 192;
constexpr int P_SHIFT = 16;                // P areas start 16 columns into the owner's S range: S columns 0..15 of a block survive the softmax
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr int Q_BYTES = 128 * 128;
constexpr int QAUG_BYTES = 128 * 32;
constexpr int KV_BYTES = KBMAX * 128;
constexpr int KAUG_BYTES = KBMAX * 32;
constexpr int OFF_Q = 0;
constexpr int OFF_K = Q_BYTES;
constexpr int OFF_V = OFF_K + KV_BYTES;
constexpr int OFF_QAUG = OFF_V + KV_BYTES;
constexpr int OFF_KAUG = OFF_QAUG + QAUG_BYTES;
constexpr int GROUP_BYTES = ((OFF_KAUG + KAUG_BYTES + 1023) / 1024) * 1024;
// Output staging: the normalised bf16 O tile of an item (128 rows x 128 B, 128-byte swizzle) leaves through ONE bulk tensor
// store per item.  Stored straight from registers, a lane's 64 bytes go to a different 128-byte line than its neighbours'
// (token rows are 2 D bytes apart): 32 memory transactions per store instruction, and the timeline showed the eight warps of a
// group queueing on the load-store unit for ~1,000 clk per item (profiles/r02_attention_trace.txt, "done" -> next "top").
#ifndef VTC_ACS_TMA_OUT
#define VTC_ACS_TMA_OUT 1
#endif
constexpr bool OUT_TMA = VTC_ACS_TMA_OUT != 0;
constexpr int OST_BYTES = 128 * 128;
constexpr int OFF_OST = GROUPS * GROUP_BYTES;                  // [GROUPS][OST_BYTES]
constexpr int OFF_CLS = OFF_OST + GROUPS * OST_BYTES;          // [GROUPS][NMAX] floats: raw logits of the CLS row
constexpr int OFF_XCH = OFF_CLS + GROUPS * NMAX * 4;           // [GROUPS][PARTS][128 rows] float2 (m, sum)
constexpr int OFF_BAR = OFF_XCH + GROUPS * PARTS * 128 * 8;
constexpr int BARS_PER_GROUP = 11;
// Experiment kept in the tree (off by default: measured no gain): -DVTC_ACS_EARLY_QK=1 issues the score MMA of the
// next item in two pieces so that it does not wait for the epilogue of the current one.  Only S columns >= O_COL overlap the O
// accumulator: keys [0, O_COL) go straight behind P V of the previous item (in-order tensor pipe), the <= 16 keys beyond once
// the softmax warps have read O out (second barrier s_tail, waited on only before a warp's chunk that starts at O_COL).
#ifndef VTC_ACS_EARLY_QK
#define VTC_ACS_EARLY_QK 0
#endif
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && KV_BYTES % 1024 == 0 && GROUP_BYTES % 1024 == 0, "swizzle atoms need 1024-byte tiles");
static_assert(SMEM_BYTES <= 232448, "attention_cs smem budget");
static_assert(KBLONG <= O_COL && KBMAX + 16 <= REGION_COLS, "TMEM region layout");

struct Params {
    const float* key_bias;   // [B,N] or null
    const uint8_t* aug;      // precomputed mask operands of every image (ops.h: AugLayout) or null: then the producer builds them from key_bias
    size_t aug_per_image;
    __nv_bfloat16* out;      // [B,N,H*64]
    float* cls_rows;         // [B,H,N] or null
    // Head-mean support ("packed P"): the un-normalised bf16 exponentials E the P V product consumes and 1 / rowsum;
    // head_mean_packed() reduces them over the heads, [B,H,N,N] fp32 is never written.
    //  * one key block (N <= 208): E is copied out of TMEM once the row is complete (P V has retired; every chunk is
    //    relative to the final maximum by then): P[b,h,r,k] = einv[b,h,r] * E[b,h,r,k]; mtab / mfin are not used;
    //  * several key blocks: a block's P is overwritten by the next block's scores, so every 32-key chunk is stored as it is
    //    produced, together with the running maximum it was taken against (the maximum may rise later; the copies in TMEM
    //    are rescaled then, the stored ones keep their own reference), and the final maximum follows per row:
    //    P[b,h,r,k] = einv[b,h,r] * 2^(mtab[b,h,r,k/32] - mfin[b,h,r]) * E[b,h,r,k].
    __nv_bfloat16* edump;    // [B,H,N,lde] or null
    float* mtab;             // [B,H,N,lde/32]
    float* mfin;             // [B,H,N]
    float* einv;             // [B,H,N]
    int lde;                 // keys per stored row: nb * KB rounded up to 32
    int B, N, H;
    int KB, nb;
    float scale, scale_log2;
    int reverse;
    unsigned long long* trace;   // debug: [grid][64 steps][SM_WARPS + GROUPS][8] %globaltimer stamps, normally null
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32b_x32(taddr, r); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }

// first chunk of column range `part` when a block has nch chunks
__device__ __forceinline__ int part_begin(int part, int nch) { return (part * nch) / PARTS; }

// 2^(-d), d >= 0 integer: exact
__device__ __forceinline__ float pow2_neg(int d) { return d >= 127 ? 0.f : __int_as_float((127 - d) << 23); }

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

// Experiment kept in the tree (off by default: measured no gain at 25-31 %, a loss at 50 %): exponentials on the
// FMA pipe.  Every VTC_ACS_POLY-th PAIR of a full chunk takes 2^x = 2^n * p(r) (n = round(x), |r| <= 1/2, p = degree-3 minimax
// fit of 2^r: relative error 9.6e-5, a twentieth of the bf16 rounding P gets anyway; 5 issue slots per element) instead of
// MUFU.EX2 (8 clk per warp instruction).  0 = every exponential on MUFU.
#ifndef VTC_ACS_POLY
#define VTC_ACS_POLY 0
#endif
#ifndef VTC_ACS_POLY_SCALAR
#define VTC_ACS_POLY_SCALAR 0
#endif
__device__ __forceinline__ float ex2_poly_scalar(float x) {      // the same on scalar instructions with immediate operands
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;
    const float r = x - (t - 12582912.0f);
    float q = fmaf(0.054526202380657196f, r, 0.2427794337272644f);
    q = fmaf(q, r, 0.6933389902114868f);
    q = fmaf(q, r, 0.9999109506607056f);
    return __uint_as_float(__float_as_uint(q) + (__float_as_uint(t) << 23));
}
__device__ __forceinline__ void ex2_poly_pair(uint64_t x2, float& e0, float& e1) {
    float x0, x1;
    unpack2(x2, x0, x1);
    if (VTC_ACS_POLY_SCALAR) {
        e0 = ex2_poly_scalar(x0);
        e1 = ex2_poly_scalar(x1);
        return;
    }
    x2 = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));        // masked keys sit at ~ -144: keep the exponent field in range
    const uint64_t t2 = add2(x2, pack2(12582912.0f, 12582912.0f));          // 1.5 * 2^23: the low mantissa bits of t are n
    const uint64_t n2 = add2(t2, pack2(-12582912.0f, -12582912.0f));
    const uint64_t r2 = fma2(n2, pack2(-1.0f, -1.0f), x2);
    uint64_t q = fma2(pack2(0.054526202380657196f, 0.054526202380657196f), r2, pack2(0.2427794337272644f, 0.2427794337272644f));
    q = fma2(q, r2, pack2(0.6933389902114868f, 0.6933389902114868f));
    q = fma2(q, r2, pack2(0.9999109506607056f, 0.9999109506607056f));
    float q0, q1, t0, t1;
    unpack2(q, q0, q1);
    unpack2(t2, t0, t1);
    e0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    e1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// One 32-key chunk: cur = raw accumulator values of this thread's row.  t_p = TMEM address of this warp's P area of the
// block, pc = index of the chunk inside that area (chunks [0, pc) are already stored there).
template <bool DUMP, bool SHORT_TAIL>      // DUMP: store the chunk of the packed P as it is produced (multi-block sequences)
__device__ __forceinline__ void chunk(const uint32_t (&cur)[32], int pc, int nvalid, float sc, float& m, uint64_t& sum2, uint32_t t_p,
                                      float* cls_dst, bool cls_thread, __nv_bfloat16* edst, float* mdst) {
    float mc = -INFINITY;
    if (nvalid >= 32) {
        if constexpr ((ABLATE & 32) != 0) mc = __uint_as_float(cur[0]);
        else
#pragma unroll
        for (int j = 0; j < 32; j += 2) mc = fmaxf(mc, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
    } else if (SHORT_TAIL && nvalid <= 8) {
        // short tail (197 tokens: 5 keys): eight columns, branch-free (the generic partial-chunk path below costs as much as a
        // full chunk, and this chunk sits on the critical path of the column part that owns it)
#pragma unroll
        for (int j = 0; j < 8; ++j) mc = fmaxf(mc, j < nvalid ? __uint_as_float(cur[j]) : -INFINITY);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) mc = fmaxf(mc, __uint_as_float(cur[j]));
    }
    const float over = fmaf(mc, sc, -m);
    const bool need = over > RESCALE_THRESHOLD;
    if (__any_sync(0xffffffffu, need)) {
        const int d = need ? static_cast<int>(ceilf(over)) : 0;
        const float f = pow2_neg(d);
        m += static_cast<float>(d);
        sum2 = mul2(sum2, pack2(f, f));
        tmem_st_wait();
        for (int cc = 0; cc < pc; ++cc) rescale_p16(t_p + cc * 16, f);
    }
    const float negm = -m;
    const uint64_t sc2 = pack2(sc, sc), negm2 = pack2(negm, negm);
    uint32_t pk[16];
    if (nvalid >= 32) {
        // all 32 exponentials are issued back to back before anything consumes them (the MUFU pipe is the bound of this loop)
        float e[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            float x0, x1;
            const uint64_t x2 = fma2(pack2u(cur[j], cur[j + 1]), sc2, negm2);
            if (VTC_ACS_POLY > 0 && ((j >> 1) % (VTC_ACS_POLY > 0 ? VTC_ACS_POLY : 1)) == (VTC_ACS_POLY > 0 ? VTC_ACS_POLY : 1) - 1) {
                ex2_poly_pair(x2, e[j], e[j + 1]);
                continue;
            }
            unpack2(x2, x0, x1);
            if constexpr ((ABLATE & 16) != 0) {
                e[j] = x0;
                e[j + 1] = x1;
                continue;
            }
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
    } else if (SHORT_TAIL && nvalid <= 8) {
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = j < nvalid ? ex2_approx(fmaf(__uint_as_float(cur[j]), sc, negm)) : 0.f;
        // pairs added one after the other, exactly like the generic partial-chunk path: the kernels with and without this path
        // (plain / packed-P variants) stay bit-identical
        sum2 = add2(add2(add2(add2(sum2, pack2(e[0], e[1])), pack2(e[2], e[3])), pack2(e[4], e[5])), pack2(e[6], e[7]));
#pragma unroll
        for (int j = 0; j < 4; ++j) pk[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
#pragma unroll
        for (int j = 4; j < 16; ++j) pk[j] = 0u;
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            if (j >= nvalid) {                        // padding keys: exact zeros, no exponentials
                pk[j >> 1] = 0u;
                continue;
            }
            float x0, x1;
            unpack2(fma2(pack2u(cur[j], cur[j + 1]), sc2, negm2), x0, x1);
            const float e0 = ex2_approx(x0);
            const float e1 = (j + 1 < nvalid) ? ex2_approx(x1) : 0.f;
            sum2 = add2(sum2, pack2(e0, e1));
            pk[j >> 1] = pack_bf16x2(e0, e1);
        }
    }
    tmem_st_32x32b_x16(t_p + pc * 16, pk);
    if constexpr (DUMP) {
        if (edst != nullptr) {                        // packed P: 64 contiguous bytes of this row + the chunk's reference maximum
            st_u8(edst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            st_u8(edst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
            *mdst = m;
        }
    }
    if (cls_thread) {                                 // raw logits of the CLS row; normalised once the row is complete
        // eight 16-byte stores (entries past nvalid are never read back)
#pragma unroll
#pragma unroll
        for (int j = 0; j < 32; j += 4) st_u4(cls_dst + j, make_uint4(cur[j], cur[j + 1], cur[j + 2], cur[j + 3]));
    }
}

// DUMP: the instantiation that also stores the packed P (kept apart: the extra pointers cost the plain kernel, which sits at
// its register cap, 6 % when they were a run-time option)
// SINGLE: the whole sequence is one key block (nb == 1, up to 208 tokens): the block loops and the cross-block rescale of O
// fold away at compile time.
template <bool DUMP, bool SINGLE>
__global__ void __launch_bounds__(THREADS, 1)
attention_cs_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmO,
                    const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    // The single-block packed-P instantiation keeps the (inert: p.trace is null) stamp code in every build: at the register cap its
    // allocation comes out better WITH it (attention + head mean 212 vs 221 us), while the other three gain 1.5-2.7 % without.
    constexpr bool TRACE = TRACE_ALL || (DUMP && SINGLE);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // group of this warp: softmax warps 0..15 by eights, then producers 16, 17 and MMA issuers 18, 19
    const int g = (warp < SM_WARPS) ? warp / GROUP_WARPS : (warp - SM_WARPS) & 1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + g * BARS_PER_GROUP;
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;
    uint64_t* k_empty = bars + 3;
    uint64_t* v_full = bars + 4;
    uint64_t* v_empty = bars + 5;
    uint64_t* s_full = bars + 6;     // S(s) ready
    uint64_t* p_full = bars + 7;     // P(s) stored by all softmax warps of the group
    uint64_t* o_full = bars + 8;     // O of an item complete
    uint64_t* o_empty = bars + 9;    // O of an item read out (all softmax warps of the group)
    uint64_t* s_tail = bars + 10;    // EARLY_QK: S columns [O_COL, nmma) of an item ready
    constexpr bool EARLY_QK = (VTC_ACS_EARLY_QK != 0) && !(DUMP && SINGLE);      // the single-block packed-P epilogue still reads P from the S columns
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_BAR + GROUPS * BARS_PER_GROUP * 8);
    uint8_t* gsm = smem + g * GROUP_BYTES;

    const int N = p.N, H = p.H, KB = p.KB, nb = SINGLE ? 1 : p.nb;
    const int qtiles = (N + 127) >> 7;
    const int n_items = p.B * H * qtiles;
    // items of this (CTA, group): global index (GROUPS * blockIdx + g) + i * GROUPS * gridDim
    const int first = GROUPS * static_cast<int>(blockIdx.x) + g, stride = GROUPS * static_cast<int>(gridDim.x);
    const int my_items = first < n_items ? (n_items - first + stride - 1) / stride : 0;
    const bool has_bias = p.key_bias != nullptr;
#ifndef VTC_ACS_NO_PREAUG
#define VTC_ACS_NO_PREAUG 0
#endif
    const bool pre_aug = (VTC_ACS_NO_PREAUG == 0) && has_bias && p.aug != nullptr;      // (the macro: A/B of the producer code, tools/build_ablate.sh)
    const int D = H * HD;
    const uint32_t kv_bytes = static_cast<uint32_t>(KB) * 128u;
    // With an even number of query tiles per (image, head) the tile index of `first + i * stride` would be the same for every
    // i, i.e. at 197 tokens group 0 would own all the full tiles (128 rows + the CLS row duty) and group 1 all the 69-row
    // ones: the two groups of a CTA swap the members of their item pair on odd iterations instead.
    const int pair_swap = (qtiles & 1) ? 0 : 1;
    auto decode = [&](int i, int& b, int& h, int& qt) {
        int it = (first + i * stride) ^ (i & pair_swap);
        if (p.reverse) it = n_items - 1 - it;
        qt = it % qtiles;
        const int bh = it / qtiles;
        b = bh / H;
        h = bh - b * H;
    };

    if (warp == SM_WARPS + GROUPS) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            tma_prefetch_desc(&tmO);
            uint64_t* all = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
            for (int gg = 0; gg < GROUPS; ++gg)
                for (int i = 0; i < BARS_PER_GROUP; ++i) mbar_init(all + gg * BARS_PER_GROUP + i, (i == 7 || i == 9) ? GROUP_WARPS : 1);      // p_full, o_empty: one arrival per softmax warp
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr + g * REGION_COLS;

    if (warp >= SM_WARPS && warp < SM_WARPS + GROUPS) {
        // ---------------- producer ----------------
        const float inv_scale = 1.0f / p.scale;
        uint32_t step = 0;
        for (int i = 0; i < my_items; ++i) {
            int b, h, qt;
            decode(i, b, h, qt);
            const float* kb = has_bias ? p.key_bias + static_cast<size_t>(b) * N : nullptr;
            if (lane == 0) mbar_wait_fast(q_empty, (i & 1) ^ 1);
            __syncwarp();
            if (has_bias && !pre_aug) {
                // no-swizzle K-major core matrices: 8 rows x 16 B, LBO (K direction) 128 B, SBO (8-row groups) 256 B
                uint8_t* qa = gsm + OFF_QAUG;
                for (int r = lane; r < 128; r += 32) {
                    const int row = qt * 128 + r;
                    const float flag = (row < N && kb[row] == 0.f) ? 1.0f : 0.0f;
                    uint8_t* dq = qa + (r >> 3) * 256 + (r & 7) * 16;
                    st_u4(dq, make_uint4(pack_bf16x2(flag, 0.f), 0u, 0u, 0u));
                    st_u4(dq + 128, make_uint4(0u, 0u, 0u, 0u));
                }
                fence_proxy_async_smem();
                __syncwarp();
            }
            if (lane == 0) {
                mbar_arrive_expect_tx(q_full, Q_BYTES + (pre_aug ? QAUG_BYTES : 0));
                tma_load_3d(gsm + OFF_Q, &tmQ, q_full, h * HD, qt * 128, b);
                if (pre_aug)      // the image's precomputed Q_aug tile (written once per layer by cls_stat_mask): a bulk copy, no generic stores
                    bulk_load_1d(gsm + OFF_QAUG, p.aug + static_cast<size_t>(b) * p.aug_per_image + static_cast<size_t>(nb) * KB * 32 + static_cast<size_t>(qt) * QAUG_BYTES,
                                 QAUG_BYTES, q_full);
            }
            for (int j = 0; j < nb; ++j, ++step) {
                const uint32_t ph = step & 1;
                if (lane == 0) mbar_wait_fast(k_empty, ph ^ 1);
                __syncwarp();
                if (has_bias && !pre_aug) {
                    uint8_t* ka = gsm + OFF_KAUG;
                    for (int r = lane; r < KB; r += 32) {
                        const int key = j * KB + r;
                        const float v = (key < N) ? kb[key] * inv_scale : 0.f;
                        uint8_t* dk = ka + (r >> 3) * 256 + (r & 7) * 16;
                        st_u4(dk, make_uint4(pack_bf16x2(v, 0.f), 0u, 0u, 0u));
                        st_u4(dk + 128, make_uint4(0u, 0u, 0u, 0u));
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                }
                if (lane == 0) {
                    unsigned long long* tr = (TRACE && p.trace && step < 64) ? p.trace + ((static_cast<size_t>(blockIdx.x) * 64 + step) * (SM_WARPS + GROUPS) + SM_WARPS + g) * 8 : nullptr;
                    mbar_arrive_expect_tx(k_full, kv_bytes + (pre_aug ? static_cast<uint32_t>(KB) * 32u : 0u));
                    tma_load_3d(gsm + OFF_K, &tmKV, k_full, D + h * HD, j * KB, b);
                    if (pre_aug)
                        bulk_load_1d(gsm + OFF_KAUG, p.aug + static_cast<size_t>(b) * p.aug_per_image + static_cast<size_t>(j) * KB * 32, static_cast<uint32_t>(KB) * 32u, k_full);
                    if (tr) { unsigned long long tt; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt)); tr[6] = tt; }      // debug: K load issued
                    mbar_wait_fast(v_empty, ph ^ 1);
                    mbar_arrive_expect_tx(v_full, kv_bytes);
                    tma_load_3d(gsm + OFF_V, &tmKV, v_full, 2 * D + h * HD, j * KB, b);
                    if (tr) { unsigned long long tt; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt)); tr[7] = tt; }      // debug: V load issued
                }
                __syncwarp();
            }
        }
    } else if (warp >= SM_WARPS + GROUPS) {
        // ---------------- MMA issuer: QK(s) -> [softmax] -> PV(s) -> QK(s+1) ... ----------------
        if (lane == 0) {
            const uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
            const uint32_t q_addr = smem_u32(gsm + OFF_Q), k_addr = smem_u32(gsm + OFF_K), v_addr = smem_u32(gsm + OFF_V);
            const uint32_t qa_addr = smem_u32(gsm + OFF_QAUG), ka_addr = smem_u32(gsm + OFF_KAUG);
            if (g == 1 && static_cast<int>(blockIdx.x) * GROUPS < n_items) {
                // start the two groups out of phase: group 1 takes off when group 0 has finished its first softmax
                uint64_t* p_full0 = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + 7;
                mbar_wait_fast(p_full0, 0);
            }
            uint32_t step = 0;
            for (int i = 0; i < my_items; ++i) {
                mbar_wait_fast(q_full, i & 1);
                for (int j = 0; j < nb; ++j, ++step) {
                    const uint32_t ph = step & 1;
                    const int vj = min(KB, N - j * KB);
                    const int nmma = (vj + 15) & ~15;
                    const int nch = (vj + 31) >> 5;
                    // the S columns are free: PV(s-1) was issued by this thread (in-order pipe).  In the single-block layout a new
                    // item's S also covers the O columns, which the softmax warps must have read out first
                    unsigned long long* tr = (TRACE && p.trace && step < 64) ? p.trace + ((static_cast<size_t>(blockIdx.x) * 64 + step) * (SM_WARPS + GROUPS) + SM_WARPS + g) * 8 : nullptr;
                    auto stamp = [&](int slot) {
                        if (tr) { unsigned long long tt; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt)); tr[slot] = tt; }
                    };
                    stamp(0);
                    if (!EARLY_QK && j == 0 && i > 0) mbar_wait_fast(o_empty, (i - 1) & 1);
                    stamp(1);
                    mbar_wait_fast(k_full, ph);
                    tc_fence_after();
                    stamp(2);
                    const int n_main = (EARLY_QK && nmma > O_COL) ? O_COL : nmma;
                    const uint32_t idesc_s = make_idesc_bf16(128, n_main, 0, 0);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(tmem_base, make_smem_desc_sw128(q_addr + k * 32, 1024, 16), make_smem_desc_sw128(k_addr + k * 32, 1024, 16), idesc_s,
                                  k != 0 ? 1u : 0u);
                    if (has_bias) umma_bf16(tmem_base, make_smem_desc(qa_addr, 256, 128, 0), make_smem_desc(ka_addr, 256, 128, 0), idesc_s, 1u);
                    umma_commit(s_full);
                    if constexpr (EARLY_QK) {
                        // O of the previous item must have been read out before S reaches into its columns and before P V overwrites it
                        if (j == 0 && i > 0) mbar_wait_fast(o_empty, (i - 1) & 1);
                        if (n_main < nmma) {
                            const uint32_t idesc_t = make_idesc_bf16(128, nmma - O_COL, 0, 0);
                            const uint32_t k_tail = k_addr + O_COL * 128;          // key row O_COL: a whole number of 8-row swizzle atoms
#pragma unroll
                            for (int k = 0; k < HD / 16; ++k)
                                umma_bf16(tmem_base + O_COL, make_smem_desc_sw128(q_addr + k * 32, 1024, 16), make_smem_desc_sw128(k_tail + k * 32, 1024, 16),
                                          idesc_t, k != 0 ? 1u : 0u);
                            if (has_bias) umma_bf16(tmem_base + O_COL, make_smem_desc(qa_addr, 256, 128, 0), make_smem_desc(ka_addr + (O_COL >> 3) * 256, 256, 128, 0), idesc_t, 1u);
                            umma_commit(s_tail);
                        }
                    }
                    umma_commit(k_empty);
                    if (j == nb - 1) umma_commit(q_empty);
                    // ---- O (+)= P(s) V(s)
                    stamp(3);
                    mbar_wait_fast(p_full, ph);
                    stamp(4);
                    mbar_wait_fast(v_full, ph);
                    tc_fence_after();
                    stamp(5);
                    const int ksteps = (ABLATE & 4) ? 1 : (nmma >> 4);
                    // every part packs its bf16 P from (16 columns past) its own first S column on: chunk c of the part that starts at
                    // chunk pb sits at column 32 pb + 16 (c - pb) + 16, i.e. k-step ks reads column 16 + 8 ks + 16 pb
                    static_assert(PARTS == 2, "the P address below assumes two column parts");
                    const int ks1 = 2 * part_begin(1, nch);                // first k-step of part 1
                    const uint64_t vdesc = make_smem_desc_sw128(v_addr, 1024, 1024);
                    for (int ks = 0; ks < ksteps; ++ks)
                        umma_bf16_ts(tmem_base + O_COL, tmem_base + P_SHIFT + 8 * ks + (ks >= ks1 ? 8 * ks1 : 0), vdesc + static_cast<uint64_t>(ks * (2048 >> 4)),
                                     idesc_o, (j | ks) != 0 ? 1u : 0u);
                    umma_commit(v_empty);
                    if (j == nb - 1) umma_commit(o_full);
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------- softmax warps ----------------
        const int part = (warp >> 2) % PARTS;
        const int quarter = warp & 3;
        const int r_local = quarter * 32 + lane;
        const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float sc = p.scale_log2;
        float2* xch = reinterpret_cast<float2*>(smem + OFF_XCH) + g * PARTS * 128;
        float* cls_buf = reinterpret_cast<float*>(smem + OFF_CLS) + g * NMAX;
        const uint32_t row_bar = 1 + g * 4 + quarter;           // named barrier of the PARTS warps that share these rows
        int s = 0;
        for (int i = 0; i < my_items; ++i) {
            int b, h, qt;
            decode(i, b, h, qt);
            const int row = qt * 128 + r_local;
            const bool warp_active = (qt * 128 + quarter * 32) < N;
            const bool cls_warp = (qt == 0) && (quarter == 0) && (p.cls_rows != nullptr);
            const bool cls_thread = cls_warp && lane == 0;
            float m = 0.f;
            uint64_t sum2 = pack2(0.f, 0.f);
            float inv = 0.f;
            constexpr bool DUMP_LOOP = DUMP && !SINGLE;
            // (the single-block variant forms its row pointer in the epilogue: nothing extra stays live across the softmax loop)
            const size_t erow_idx = DUMP_LOOP ? (static_cast<size_t>(b) * H + h) * N + row : 0;
            __nv_bfloat16* erow = (DUMP_LOOP && row < N) ? p.edump + erow_idx * p.lde : nullptr;
            float* mrow = (DUMP_LOOP && row < N) ? p.mtab + erow_idx * (p.lde >> 5) : nullptr;
            for (int j = 0; j < nb; ++j, ++s) {
                const int vj = min(KB, N - j * KB);
                const int nch = (vj + 31) >> 5;
                const int c0 = part_begin(part, nch), c1 = part_begin(part + 1, nch);      // my 32-key chunks of this block
                const uint32_t t_p = t_s + c0 * 32 + P_SHIFT;                              // my P area: on top of S columns I have consumed
                unsigned long long* tr = (TRACE && p.trace && s < 64 && lane == 0) ? p.trace + ((static_cast<size_t>(blockIdx.x) * 64 + s) * (SM_WARPS + GROUPS) + warp) * 8 : nullptr;
                auto stamp = [&](int slot) {
                    if (tr) { unsigned long long tt; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt)); tr[slot] = tt; }
                };
                stamp(0);
                mbar_wait_fast(s_full, s & 1);
                tc_fence_after();
                stamp(1);
                if (warp_active) {
                    if (j == 0) {
                        uint32_t f8[8];
                        tmem_ld_32x32b_x8(t_s, f8);              // S columns 0..7 are never overwritten by P (P_SHIFT)
                        tmem_ld_wait();
                        float m0 = __uint_as_float(f8[0]);            // key 0 (CLS) always exists and is never masked
#pragma unroll
                        for (int q = 1; q < 8; ++q)
                            if (q < vj) m0 = fmaxf(m0, __uint_as_float(f8[q]));
                        m = ceilf(m0 * sc);
                    }
                    const float m_start = m;
                    // tcgen05.ld stalls this warp for the whole transfer, so there is nothing to gain from prefetching into a second
                    // register buffer: the other warps of the sub-partition fill the time
                    for (int c = c0; c < c1; ++c) {
                        uint32_t cur[32];
                        const int nvalid = vj - c * 32;
                        if constexpr (EARLY_QK) {
                            if (c * 32 >= O_COL) {          // these columns come from the second piece of Q K^T (one per item)
                                mbar_wait_fast(s_tail, i & 1);
                                tc_fence_after();
                            }
                        }
                        if constexpr ((ABLATE & 2) != 0) {
#pragma unroll
                            for (int q = 0; q < 32; ++q) cur[q] = __float_as_uint(static_cast<float>(q + lane) * 0.01f);
                        } else {
                            if (nvalid > 16) tmem_ld_32x32b_x32(t_s + c * 32, cur);
                            else if (nvalid > 8 || DUMP) tmem_ld_32x32b_x16(t_s + c * 32, reinterpret_cast<uint32_t(&)[16]>(cur));
                            else tmem_ld_32x32b_x8(t_s + c * 32, reinterpret_cast<uint32_t(&)[8]>(cur));
                            tmem_ld_wait();
                        }
                        if constexpr ((ABLATE & 1) != 0) {
                            uint32_t pk0[16];
#pragma unroll
                            for (int q = 0; q < 16; ++q) pk0[q] = cur[q] & 0x3f803f80u;
                            tmem_st_32x32b_x16(t_p + (c - c0) * 16, pk0);
                            sum2 = add2(sum2, pack2(1.0f, 1.0f));
                        } else {
                            chunk<DUMP_LOOP, !DUMP>(cur, c - c0, nvalid, sc, m, sum2, t_p, cls_buf + j * KB + c * 32, cls_thread,
                                             (DUMP_LOOP && erow) ? erow + j * KB + c * 32 : nullptr, (DUMP_LOOP && mrow) ? mrow + ((j * KB) >> 5) + c : nullptr);
                        }
                    }
                    stamp(2);
                    // ---- the warps that share these rows agree on the row maximum (and, at the end, on the row sum)
                    float s0, s1;
                    unpack2(sum2, s0, s1);
                    // (no barrier before the write: whoever still has to READ the previous block's exchange slots does so before it
                    // arrives on p_full, and this warp only got here through s_full / o_full of a later MMA, which waited for all
                    // eight arrivals)
                    xch[part * 128 + r_local] = make_float2(m, s0 + s1);
                    named_bar_sync(row_bar, 32 * PARTS);
                    stamp(3);
                    float2 other[PARTS];
                    float m_blk = m;
#pragma unroll
                    for (int q = 0; q < PARTS; ++q) {
                        other[q] = xch[q * 128 + r_local];
                        m_blk = fmaxf(m_blk, other[q].x);
                    }
                    if (__any_sync(0xffffffffu, m < m_blk)) {       // rare: another part raised the maximum further than I did
                        const float f = pow2_neg(static_cast<int>(m_blk - m));
                        sum2 = mul2(sum2, pack2(f, f));
                        tmem_st_wait();
                        for (int cc = 0; cc < c1 - c0; ++cc) rescale_p16(t_p + cc * 16, f);
                        m = m_blk;
                    }
                    if (j > 0 && __any_sync(0xffffffffu, m_blk > m_start)) {      // rare: earlier key blocks were accumulated against a lower m
                        // PV(s-1) has retired: S(s) was committed behind it on the in-order tensor pipe
                        const float f = pow2_neg(static_cast<int>(m_blk - m_start));
#pragma unroll
                        for (int q = 0; q < OCOLS / 16; ++q) rescale_f16(t_s + O_COL + part * OCOLS + q * 16, f);
                    }
                    tmem_st_wait();
                    if (j == nb - 1) {
                        float total = 0.f;
#pragma unroll
                        for (int q = 0; q < PARTS; ++q) total += other[q].y * pow2_neg(static_cast<int>(m_blk - other[q].x));
                        inv = 1.0f / total;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                stamp(4);
            }
            if (cls_warp && part == 0) {
                // P[b,h,0,:] = 2^(x - m) / rowsum from the staged logits of all parts (visible after the row barrier)
                const float m0 = __shfl_sync(0xffffffffu, m, 0), inv0 = __shfl_sync(0xffffffffu, inv, 0);
                float* dst = p.cls_rows + (static_cast<size_t>(b) * H + h) * N;
                for (int jj = lane; jj < N; jj += 32) dst[jj] = ex2_approx(fmaf(cls_buf[jj], sc, -m0)) * inv0;
            }
            // ---- epilogue: O / rowsum -> bf16 (the other group owns the TMEM port meanwhile)
            unsigned long long* tre = (TRACE && p.trace && s - 1 < 64 && lane == 0) ? p.trace + ((static_cast<size_t>(blockIdx.x) * 64 + (s - 1)) * (SM_WARPS + GROUPS) + warp) * 8 : nullptr;
            mbar_wait_fast(o_full, i & 1);
            if (tre) { unsigned long long tt; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt)); tre[5] = tt; }
            tc_fence_after();
            if (warp_active) {
                uint32_t o[OCOLS];
                if constexpr ((ABLATE & 8) != 0) {
#pragma unroll
                    for (int q = 0; q < OCOLS; ++q) o[q] = __float_as_uint(static_cast<float>(q + lane));
                } else {
                    tmem_ld_cols(t_s + O_COL + part * OCOLS, o);
                    tmem_ld_wait();
                }
                if constexpr (DUMP && SINGLE) {
                    // the bf16 exponentials of my column range are still in TMEM (P V has retired): one 64-byte row segment per
                    // 32-key chunk, two full 32-byte sectors per lane
                    const int nch = (N + 31) >> 5;
                    const int c0 = part_begin(part, nch), c1 = part_begin(part + 1, nch);
                    const size_t ridx = (static_cast<size_t>(b) * H + h) * N + (row < N ? row : 0);
                    __nv_bfloat16* er = p.edump + ridx * p.lde + c0 * 32;
                    for (int pc = 0; pc < c1 - c0; ++pc) {
                        uint32_t q[16];
                        tmem_ld_32x32b_x16(t_s + c0 * 32 + P_SHIFT + pc * 16, q);
                        tmem_ld_wait();
                        if (row < N) {
                            st_u8(er + pc * 32, q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
                            st_u8(er + pc * 32 + 16, q[8], q[9], q[10], q[11], q[12], q[13], q[14], q[15]);
                        }
                    }
                    if (part == 0 && row < N) p.einv[ridx] = inv;
                }
                if constexpr (DUMP && !SINGLE) {
                    if (erow != nullptr && part == 0) {
                        p.einv[erow_idx] = inv;
                        p.mfin[erow_idx] = m;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty);
                if (row < N) {
                    // OUT_TMA: my 64 bytes of the staged tile row (16-byte chunk index XOR row % 8: the 128-byte swizzle of the store's
                    // tensor map; a quarter warp writes eight different chunks: no bank conflict)
                    uint8_t* srow = smem + OFF_OST + g * OST_BYTES + r_local * 128;
                    __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * N + row) * D + h * HD + part * OCOLS;
#pragma unroll
                    for (int g4 = 0; g4 < OCOLS / 8; ++g4) {
                        const uint4 v = make_uint4(pack_bf16x2(__uint_as_float(o[8 * g4]) * inv, __uint_as_float(o[8 * g4 + 1]) * inv),
                                                   pack_bf16x2(__uint_as_float(o[8 * g4 + 2]) * inv, __uint_as_float(o[8 * g4 + 3]) * inv),
                                                   pack_bf16x2(__uint_as_float(o[8 * g4 + 4]) * inv, __uint_as_float(o[8 * g4 + 5]) * inv),
                                                   pack_bf16x2(__uint_as_float(o[8 * g4 + 6]) * inv, __uint_as_float(o[8 * g4 + 7]) * inv));
                        if constexpr (OUT_TMA) st_u4(srow + (((part * (OCOLS / 8) + g4) ^ (r_local & 7)) << 4), v);
                        else st_u4(dst + 8 * g4, v);
                    }
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty);
            }
            if constexpr (OUT_TMA) {
                // all eight warps of the group have staged their rows -> one thread stores the tile (token rows >= N are clipped by
                // the tensor map).  The staging tile is reused by the next item: the storing thread waits until the store has READ it;
                // that wait is ordered before every later write by this thread's own arrival on p_full of the next item (-> P V ->
                // o_full, which every warp waits for before it gets here again).
                fence_proxy_async_smem();
                named_bar_sync(9 + g, GROUP_WARPS * 32);
                if ((warp % GROUP_WARPS) == 0 && lane == 0) {
                    tma_store_3d(&tmO, smem + OFF_OST + g * OST_BYTES, h * HD, qt * 128, b);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                }
            }
            if (cls_warp) named_bar_sync(row_bar, 32 * PARTS);     // the CLS staging buffer may be overwritten by the next item
            if (tre) { unsigned long long tt; asm volatile("mov.u64 %0, %%clock64;" : "=l"(tt)); tre[6] = tt; }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == SM_WARPS + GROUPS) tmem_dealloc(*tmem_ptr, 512);
}
}  // namespace acs

int attention_cs(const void* qkv, const float* key_bias, void* out, float* cls_rows, int batch, int n_tokens, int heads, float scale,
                 cudaStream_t stream, int reverse, const PackedP* packed, const void* aug) {
    using namespace acs;
    VTC_REQUIRE(qkv && out, VTC_ERR_ARG, "attention: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0, VTC_ERR_SHAPE, "attention: bad shape");
    VTC_REQUIRE(scale > 0.f, VTC_ERR_ARG, "attention: scale must be positive");
    VTC_REQUIRE(n_tokens <= NMAX, VTC_ERR_SHAPE, "attention: %d tokens > %d", n_tokens, NMAX);
    VTC_REQUIRE(packed == nullptr || (packed->e && packed->einv && (n_tokens <= KBMAX || (packed->mtab && packed->mfin))), VTC_ERR_ARG,
                "attention: incomplete packed-P output");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    Params p{};
    p.key_bias = key_bias;
    p.aug = key_bias ? static_cast<const uint8_t*>(aug) : nullptr;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.cls_rows = cls_rows;
    if (packed) {
        p.edump = static_cast<__nv_bfloat16*>(packed->e);
        p.mtab = packed->mtab;
        p.mfin = packed->mfin;
        p.einv = packed->einv;
    }
    p.B = batch;
    p.N = n_tokens;
    p.H = heads;
    if (n_tokens <= KBMAX) {
        p.nb = 1;
        p.KB = (n_tokens + 15) & ~15;
    } else {
        p.nb = cdiv(n_tokens, KBLONG);
        p.KB = (cdiv(n_tokens, p.nb) + 31) & ~31;
    }
    {
        const AugLayout al = attention_aug_layout(n_tokens);
        VTC_REQUIRE(al.KB == p.KB && al.nb == p.nb, VTC_ERR_SHAPE, "attention: mask-operand layout out of sync with the key blocking");
        VTC_REQUIRE(!p.aug || (reinterpret_cast<uintptr_t>(p.aug) & 15) == 0, VTC_ERR_ARG, "attention: mask operands must be 16-byte aligned");
        p.aug_per_image = al.per_image;
    }
    p.lde = attention_packed_ld(n_tokens);
    VTC_REQUIRE(p.lde == (p.nb == 1 ? ((n_tokens + 31) & ~31) : p.nb * p.KB), VTC_ERR_SHAPE, "attention: packed-P row stride out of sync with the key blocking");
    p.scale = scale;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.reverse = reverse;
    p.trace = attention_trace_buffer();
    const int D = heads * HD;
    uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)n_tokens, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)3 * D * 2, (uint64_t)n_tokens * 3 * D * 2};
    CUtensorMap tmQ, tmKV, tmO;
    uint32_t boxq[3] = {HD, 128, 1};
    uint32_t boxkv[3] = {HD, static_cast<uint32_t>(p.KB), 1};
    rc = make_tmap_bf16(&tmQ, qkv, 3, dims, strides, boxq);
    if (rc != VTC_OK) return rc;
    rc = make_tmap_bf16(&tmKV, qkv, 3, dims, strides, boxkv);
    if (rc != VTC_OK) return rc;
    uint64_t odims[3] = {(uint64_t)D, (uint64_t)n_tokens, (uint64_t)batch};
    uint64_t ostrides[2] = {(uint64_t)D * 2, (uint64_t)n_tokens * D * 2};
    rc = make_tmap_bf16(&tmO, out, 3, odims, ostrides, boxq);
    if (rc != VTC_OK) return rc;
    static SmemOptIn optin[4];
    if ((rc = optin[0].ensure(reinterpret_cast<const void*>(attention_cs_kernel<false, false>), SMEM_BYTES)) != VTC_OK) return rc;
    if ((rc = optin[1].ensure(reinterpret_cast<const void*>(attention_cs_kernel<false, true>), SMEM_BYTES)) != VTC_OK) return rc;
    if ((rc = optin[2].ensure(reinterpret_cast<const void*>(attention_cs_kernel<true, false>), SMEM_BYTES)) != VTC_OK) return rc;
    if ((rc = optin[3].ensure(reinterpret_cast<const void*>(attention_cs_kernel<true, true>), SMEM_BYTES)) != VTC_OK) return rc;
    const int items = batch * heads * cdiv(n_tokens, 128);
    int grid = cdiv(items, GROUPS);
    if (grid > device_sm_count()) grid = device_sm_count();
    const bool single = p.nb == 1;
    if (packed && single) attention_cs_kernel<true, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmO, p);
    else if (packed) attention_cs_kernel<true, false><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmO, p);
    else if (single) attention_cs_kernel<false, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmO, p);
    else attention_cs_kernel<false, false><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmO, p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

}  // namespace vtc
