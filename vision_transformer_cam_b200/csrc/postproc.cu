// CAM / attention-rollout / pseudo-label post-processing: the math predict.py:189-269 and
// validate.py:132-276 run after the forward, as bandwidth-oriented kernels (coalesced row reads, warp-shuffle
// reductions, fused upsample + argmax so that only 1 byte per output pixel reaches HBM).
#include "common.cuh"
#include "ops.h"

namespace vtc {

static int grid_cap(size_t blocks, int per_sm) {
    const size_t cap = static_cast<size_t>(device_sm_count()) * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return static_cast<int>(blocks);
}

// ---- attention rollout ----------------------------------------------------------------------------------------
// r <- r . A_l for l = L-1 .. 0, A_l = (Pbar_l + I) / rowsum(Pbar_l + I): only the CLS row of the product is consumed
// (predict.py:229-232), so the dense N^3 chain collapses to a vector-matrix chain.  One CTA per image; each warp
// streams whole rows (coalesced, four rows in flight, the values kept in registers), gets the row sums by shuffle, and
// accumulates w_i * row into its private column accumulators in shared memory; the matrix is read exactly once.  This is the
// fp32 entry point (vtc_rollout, exact on fp32 head means); the forward's own rollout runs on bf16 operands (below).
constexpr int ROLL_WARPS = 8;
constexpr int ROLL_ROWS = 4;          // rows a warp keeps in flight (the kernel is latency-bound: one row per warp at a time took 4x longer)
constexpr int ROLL_MAXV = 8;          // values per lane and row held in registers: n_tokens <= 256; longer rows take the two-pass path
__global__ void __launch_bounds__(ROLL_WARPS * 32) rollout_kernel(const float* __restrict__ pbar, float* __restrict__ out, int L, int B, int N) {
    extern __shared__ float sm[];
    float* r = sm;                  // [N]
    float* acc = sm + N;            // [ROLL_WARPS][N]
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < N; j += blockDim.x) r[j] = (j == 0) ? 1.0f : 0.0f;
    __syncthreads();
    const bool in_regs = N <= 32 * ROLL_MAXV;
    for (int l = L - 1; l >= 0; --l) {
        const float* A = pbar + (static_cast<size_t>(l) * B + b) * N * N;
        float* my = acc + warp * N;
        for (int j = lane; j < N; j += 32) my[j] = 0.f;
        if (in_regs) {
            for (int i0 = warp * ROLL_ROWS; i0 < N; i0 += ROLL_WARPS * ROLL_ROWS) {
                float v[ROLL_ROWS][ROLL_MAXV], s[ROLL_ROWS];
#pragma unroll
                for (int q = 0; q < ROLL_ROWS; ++q) {          // all loads of the four rows are issued before anything consumes them
                    const int i = i0 + q;
                    const bool live = i < N && r[i] != 0.0f;   // exact zeros only (the first step has a single non-zero row)
                    const float* row = A + static_cast<size_t>(i < N ? i : 0) * N;
#pragma unroll
                    for (int k = 0; k < ROLL_MAXV; ++k) v[q][k] = (live && lane + 32 * k < N) ? __ldg(row + lane + 32 * k) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < ROLL_ROWS; ++q) {
                    float t = 0.f;
#pragma unroll
                    for (int k = 0; k < ROLL_MAXV; ++k) t += v[q][k];          // same order as the one-row loop: bit-identical sums
                    s[q] = warp_sum(t);
                }
#pragma unroll
                for (int q = 0; q < ROLL_ROWS; ++q) {
                    const int i = i0 + q;
                    if (i >= N) break;
                    const float ri = r[i];
                    if (ri == 0.0f) continue;
                    const float w = ri / (s[q] + 1.0f);
#pragma unroll
                    for (int k = 0; k < ROLL_MAXV; ++k) {
                        const int j = lane + 32 * k;
                        if (j < N) my[j] += w * v[q][k] + (j == i ? w : 0.f);          // + identity
                    }
                }
            }
        } else {
            for (int i = warp; i < N; i += ROLL_WARPS) {
                const float ri = r[i];
                if (ri == 0.0f) continue;
                const float* row = A + static_cast<size_t>(i) * N;
                float s = 0.f;
                for (int j = lane; j < N; j += 32) s += row[j];
                const float w = ri / (warp_sum(s) + 1.0f);
                for (int j = lane; j < N; j += 32) my[j] += w * row[j] + (j == i ? w : 0.f);   // second touch hits L1; + identity
            }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < ROLL_WARPS; ++w) s += acc[w * N + j];
            r[j] = s;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x + 1; j < N; j += blockDim.x) out[static_cast<size_t>(b) * (N - 1) + j - 1] = r[j];
}

int rollout(const float* attn_mean, float* row, int layers, int batch, int n_tokens, cudaStream_t stream) {
    VTC_REQUIRE(attn_mean && row, VTC_ERR_ARG, "rollout: null pointer");
    VTC_REQUIRE(layers > 0 && batch > 0 && n_tokens > 1, VTC_ERR_SHAPE, "rollout: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const size_t smem = sizeof(float) * static_cast<size_t>(n_tokens) * (ROLL_WARPS + 1);
    VTC_REQUIRE(smem <= 48 * 1024, VTC_ERR_SHAPE, "rollout: %d tokens need %zu bytes of smem", n_tokens, smem);
    rollout_kernel<<<batch, ROLL_WARPS * 32, smem, stream>>>(attn_mean, row, layers, batch, n_tokens);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- attention rollout, streaming version ------------------------------------------------------------------------
// Same chain on the bf16 "rollout operands" (ops.h: row = N values | zero padding | fp32 row sum), the layout the forward
// keeps for the rollout: half the bytes of fp32 [B,N,N] matrices and 16-byte aligned rows, so a block of rows is ONE bulk
// copy (cp.async.bulk, no descriptor) into shared memory.  One CTA per image walks layers L-1 .. 0 and, inside a layer, row
// blocks of <= 32 KB through a 3-stage ring: while block t is folded into the running vector (thread = column pair x row
// parity, packed fp32x2 FMAs, conflict-free 4-byte shared loads), blocks t+1 and t+2 are in flight -- 64 KB per CTA, two
// CTAs per SM, more than the ~45 KB per SM that the HBM latency-bandwidth product asks for.  Only the CLS row of the last
// layer is read (the chain starts from e0).  Algorithmic bytes: (L-1) N ldr 2 + ldr 2 per image (0.87 MB at N = 197, L = 12).
namespace rollop {
constexpr int THREADS = 256;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 32 * 1024;
constexpr int MAX_PAIRS = 8;        // column pairs per thread: N <= 2048
}  // namespace rollop

__global__ void __launch_bounds__(rollop::THREADS, 2)
rollout_operand_kernel(const __nv_bfloat16* __restrict__ ops, float* __restrict__ out, int L, int B, int N, int ldr, int rows_per_block,
                       int blocks_per_layer) {
    using namespace rollop;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* stage = smem_raw;                                                    // [STAGES][STAGE_BYTES]
    float* r = reinterpret_cast<float*>(smem_raw + STAGES * STAGE_BYTES);         // [ldr]  running vector
    float* part = r + ldr;                                                        // [ldr]  odd-row partial sums of a layer
    float* w = part + ldr;                                                        // [rows_per_block] weights of the block in hand
    uint64_t* bar = reinterpret_cast<uint64_t*>(w + ((rows_per_block + 1) & ~1));
    const int b = blockIdx.x, tid = threadIdx.x;
    const int pairs = ldr >> 1;
    const int pcol = tid & 127, par = tid >> 7;            // my column pairs: pcol + 128 k; my rows of a block: par, par + 2, ...
    const int T = 1 + (L - 1) * blocks_per_layer;          // block 0 = the CLS row of layer L-1, then whole layers L-2 .. 0

    auto block_of = [&](int t, int& l, int& row0, int& nrows) {
        if (t == 0) { l = L - 1; row0 = 0; nrows = 1; return; }
        const int u = t - 1;
        l = L - 2 - u / blocks_per_layer;
        row0 = (u % blocks_per_layer) * rows_per_block;
        nrows = min(rows_per_block, N - row0);
    };
    auto issue = [&](int t) {                               // one thread
        int l, row0, nrows;
        block_of(t, l, row0, nrows);
        const uint32_t bytes = static_cast<uint32_t>(nrows) * ldr * 2;
        const __nv_bfloat16* src = ops + ((static_cast<size_t>(l) * B + b) * N + row0) * ldr;
        mbar_arrive_expect_tx(&bar[t % STAGES], bytes);
        bulk_load_1d(stage + (t % STAGES) * STAGE_BYTES, src, bytes, &bar[t % STAGES]);
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        fence_barrier_init();
        for (int t = 0; t < STAGES && t < T; ++t) issue(t);
    }
    for (int j = tid; j < ldr; j += THREADS) r[j] = (j == 0) ? 1.0f : 0.0f;
    uint64_t acc[MAX_PAIRS];
#pragma unroll
    for (int k = 0; k < MAX_PAIRS; ++k) acc[k] = 0ull;
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        int l, row0, nrows;
        block_of(t, l, row0, nrows);
        const uint8_t* data = stage + (t % STAGES) * STAGE_BYTES;
        mbar_wait(&bar[t % STAGES], (t / STAGES) & 1);
        // weights of the block's rows: w_i = r_i / rowsum(Pbar_i + I), the row sum rides in the row's last four bytes
        for (int i = tid; i < nrows; i += THREADS)
            w[i] = r[row0 + i] / (*reinterpret_cast<const float*>(data + (static_cast<size_t>(i) * ldr + ldr - 2) * 2) + 1.0f);
        __syncthreads();
        const uint32_t* d32 = reinterpret_cast<const uint32_t*>(data);
#pragma unroll
        for (int k = 0; k < MAX_PAIRS; ++k) {
            const int p = pcol + 128 * k;
            if (p < pairs) {
                uint64_t a = acc[k];
                for (int i = par; i < nrows; i += 2) {
                    const uint32_t v = d32[i * pairs + p];
                    const float wi = w[i];
                    a = fma2(pack2(wi, wi), pack2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)), a);
                }
                if (par == 0) {      // the identity of Pbar + I: column j receives w_j when row j is in this block
                    const int i0 = 2 * p - row0, i1 = 2 * p + 1 - row0;
                    a = add2(a, pack2((i0 >= 0 && i0 < nrows) ? w[i0] : 0.f, (i1 >= 0 && i1 < nrows) ? w[i1] : 0.f));
                }
                acc[k] = a;
            }
        }
        __syncthreads();                                     // the stage and w[] are free again
        if (tid == 0 && t + STAGES < T) issue(t + STAGES);
        const bool layer_done = (t == 0) || ((t - 1) % blocks_per_layer == blocks_per_layer - 1);
        if (layer_done) {
            // r_new = even-row sums + odd-row sums
            if (par == 1) {
#pragma unroll
                for (int k = 0; k < MAX_PAIRS; ++k) {
                    const int p = pcol + 128 * k;
                    if (p < pairs) { float lo, hi; unpack2(acc[k], lo, hi); part[2 * p] = lo; part[2 * p + 1] = hi; acc[k] = 0ull; }
                }
            }
            __syncthreads();
            if (par == 0) {          // every read of the old r (the w[] of this layer's blocks) is behind the barriers above
#pragma unroll
                for (int k = 0; k < MAX_PAIRS; ++k) {
                    const int p = pcol + 128 * k;
                    if (p < pairs) {
                        float lo, hi;
                        unpack2(acc[k], lo, hi);
                        acc[k] = 0ull;
                        r[2 * p] = (2 * p < N) ? lo + part[2 * p] : 0.f;
                        r[2 * p + 1] = (2 * p + 1 < N) ? hi + part[2 * p + 1] : 0.f;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid + 1; j < N; j += THREADS) out[static_cast<size_t>(b) * (N - 1) + j - 1] = r[j];
}

int rollout_operand(const void* operands, float* row, int layers, int batch, int n_tokens, cudaStream_t stream) {
    using namespace rollop;
    VTC_REQUIRE(operands && row, VTC_ERR_ARG, "rollout_operand: null pointer");
    VTC_REQUIRE(layers > 0 && batch > 0 && n_tokens > 1 && n_tokens <= 2 * 128 * MAX_PAIRS - 2, VTC_ERR_SHAPE, "rollout_operand: bad shape");
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(operands) & 15) == 0, VTC_ERR_ARG, "rollout_operand: operands must be 16-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int ldr = rollout_operand_ld(n_tokens);
    const int max_rows = STAGE_BYTES / (ldr * 2);
    const int blocks_per_layer = cdiv(n_tokens, max_rows);
    const int rows_per_block = cdiv(n_tokens, blocks_per_layer);
    const size_t smem = static_cast<size_t>(STAGES) * STAGE_BYTES + sizeof(float) * (2 * ldr + ((rows_per_block + 1) & ~1)) + 8 * STAGES;
    static SmemOptIn optin;
    if ((rc = optin.ensure(reinterpret_cast<const void*>(rollout_operand_kernel), smem)) != VTC_OK) return rc;
    rollout_operand_kernel<<<batch, THREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(operands), row, layers, batch, n_tokens, ldr,
                                                             rows_per_block, blocks_per_layer);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// fp32 [B,N,N] head mean -> rollout operand: warp per row (the forwards whose head mean comes from a full fp32 P)
__global__ void rollout_operand_from_mean_kernel(const float* __restrict__ mean, __nv_bfloat16* __restrict__ op, int rows, int N, int ldr) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int rid = warp; rid < rows; rid += nwarps) {
        const float* src = mean + static_cast<size_t>(rid) * N;
        __nv_bfloat16* dst = op + static_cast<size_t>(rid) * ldr;
        float s = 0.f;
        for (int j = lane; j < ldr - 2; j += 32) {
            const float v = j < N ? bf16_round(src[j]) : 0.f;
            s += v;
            dst[j] = __float2bfloat16_rn(v);
        }
        s = warp_sum(s);
        if (lane == 0) *reinterpret_cast<float*>(dst + ldr - 2) = s;
    }
}

int rollout_operand_from_mean(const float* mean, void* operand, int batch, int n_tokens, cudaStream_t stream) {
    VTC_REQUIRE(mean && operand, VTC_ERR_ARG, "rollout_operand_from_mean: null pointer");
    VTC_REQUIRE(batch > 0 && n_tokens > 1, VTC_ERR_SHAPE, "rollout_operand_from_mean: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int rows = batch * n_tokens;
    rollout_operand_from_mean_kernel<<<grid_cap(cdiv(rows, 8), 8), 256, 0, stream>>>(mean, static_cast<__nv_bfloat16*>(operand), rows, n_tokens,
                                                                                     rollout_operand_ld(n_tokens));
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- CLS-row maps (per-layer maps predict.py:261-266; layers-6..12 bg map validate.py:225-237) -----------------
__global__ void cls_layer_map_kernel(const float* __restrict__ cls_rows, float* __restrict__ map, int first, int last, int B, int H, int N) {
    extern __shared__ float rowv[];      // [N]
    __shared__ float red[32];
    const int b = blockIdx.x;
    const float inv = 1.0f / (static_cast<float>(last - first) * H);
    float part = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float s = 0.f;
        for (int l = first; l < last; ++l) {
            const float* src = cls_rows + ((static_cast<size_t>(l) * B + b) * H) * N + j;
            for (int h = 0; h < H; ++h) s += src[static_cast<size_t>(h) * N];
        }
        s *= inv;
        if (j == 0) s += 1.0f;
        rowv[j] = s;
        part += s;
    }
    // block reductions (sum, then max of the patch part)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    float tot = (lane < (blockDim.x >> 5)) ? red[lane] : 0.f;
    tot = warp_sum(tot);
    __syncthreads();
    float mx = 0.f;
    for (int j = threadIdx.x + 1; j < N; j += blockDim.x) {
        const float v = rowv[j] / tot;
        rowv[j] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    float gm = (lane < (blockDim.x >> 5)) ? red[lane] : 0.f;
    gm = warp_max(gm);
    for (int j = threadIdx.x + 1; j < N; j += blockDim.x) map[static_cast<size_t>(b) * (N - 1) + j - 1] = rowv[j] / gm;
}

int cls_layer_map(const float* cls_rows, float* map, int layers, int first, int last, int batch, int heads, int n_tokens, cudaStream_t stream) {
    VTC_REQUIRE(cls_rows && map, VTC_ERR_ARG, "cls_layer_map: null pointer");
    VTC_REQUIRE(0 <= first && first < last && last <= layers && batch > 0 && heads > 0 && n_tokens > 1, VTC_ERR_SHAPE, "cls_layer_map: bad range [%d,%d) of %d", first, last, layers);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    cls_layer_map_kernel<<<batch, 256, sizeof(float) * n_tokens, stream>>>(cls_rows, map, first, last, batch, heads, n_tokens);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- classic CAM ---------------------------------------------------------------------------------------------------
// cam[b,c,p] = <W[c,:], F[b,p,:]> on the block-L patch tokens, ReLU, per-map min-max (t.py:66-70, utils.py:84-85).
// A [P x D] x [D x C] product per image with C = 20: HBM-bound (602 KB of fp32 tokens per image, read once).  One CTA per
// image, one warp per 16-patch tile; the product runs on the warp-level tensor cores with fp32-equivalent operands: tokens
// and weights are split into (hi, lo) bf16 pairs on the fly (x ~= hi + lo, 16 mantissa bits) and every product is
// hi.hi + lo.hi + hi.lo with fp32 accumulation -- three mma.sync per k-step instead of 40 running dot products and 40
// shuffle reductions per lane, which made the first (FMA-pipe) version of this kernel latency-bound at 13 % of the HBM rate.
// W sits in shared memory as [C8][D+8] hi / lo (bank-conflict-free B fragments), the raw maps as [C][P] for the min-max pass.
__device__ __forceinline__ void mma_bf16_16816_acc(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pair(float2 v, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(v.x, v.y);
    lo = pack_bf16x2(v.x - __uint_as_float(hi << 16), v.y - __uint_as_float(hi & 0xffff0000u));
}

// COS: the same product with PER-IMAGE weights (w + b * w_stride: the image's K high-weight tokens) and a cosine epilogue,
// out[b,k,p] = <F_p, o_k> / (max(|F_p|, 1e-12) max(|o_k|, 1e-12)) (validate.py:157-175: F.normalize on both sides, then the
// dot product); the patch norms are accumulated from the A fragments on the way.
template <int NT, bool COS>      // class tiles of 8: C <= 8 * NT
__global__ void __launch_bounds__(512) cam_project_kernel(const float* __restrict__ tokens, const float* __restrict__ w_all, size_t w_stride,
                                                          float* __restrict__ cam, int N, int D, int C, int relu, float eps) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    __shared__ float wnorm[64];
    const int P = N - 1;
    const int wp = D + 8;                                                   // W row pitch (bf16)
    __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(sm_raw);          // [8 NT][wp]
    __nv_bfloat16* wlo = whi + static_cast<size_t>(8 * NT) * wp;
    float* raw = reinterpret_cast<float*>(wlo + static_cast<size_t>(8 * NT) * wp);      // [C][P]
    const int b = blockIdx.x;
    const float* w = w_all + static_cast<size_t>(b) * w_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    if (COS) {
        for (int c = warp; c < C; c += nwarps) {
            float s = 0.f;
            for (int d = lane; d < D; d += 32) { const float v = __ldg(w + static_cast<size_t>(c) * D + d); s = fmaf(v, v, s); }
            s = warp_sum(s);
            if (lane == 0) wnorm[c] = fmaxf(sqrtf(s), 1e-12f);            // F.normalize eps
        }
    }
    for (int i = threadIdx.x; i < 8 * NT * (D / 2); i += blockDim.x) {
        const int c = i / (D / 2), d2 = i - c * (D / 2);
        uint32_t hi = 0u, lo = 0u;
        if (c < C) split_pair(__ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(c) * D) + d2), hi, lo);
        *reinterpret_cast<uint32_t*>(whi + static_cast<size_t>(c) * wp + 2 * d2) = hi;
        *reinterpret_cast<uint32_t*>(wlo + static_cast<size_t>(c) * wp + 2 * d2) = lo;
    }
    __syncthreads();
    const float* F = tokens + (static_cast<size_t>(b) * N + 1) * D;
    for (int tile = warp; tile * 16 < P; tile += nwarps) {
        const int p0 = tile * 16 + g, p1 = p0 + 8;
        const float* f0 = F + static_cast<size_t>(p0 < P ? p0 : P - 1) * D + 2 * t;
        const float* f1 = F + static_cast<size_t>(p1 < P ? p1 : P - 1) * D + 2 * t;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        float nn0 = 0.f, nn1 = 0.f;                     // COS: squared norms of patch rows p0 / p1 (this lane's 4 of every 16 columns)
#pragma unroll 4
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t ah[4], al[4];
            const float2 a00 = *reinterpret_cast<const float2*>(f0 + ks * 16), a10 = *reinterpret_cast<const float2*>(f1 + ks * 16);
            const float2 a01 = *reinterpret_cast<const float2*>(f0 + ks * 16 + 8), a11 = *reinterpret_cast<const float2*>(f1 + ks * 16 + 8);
            if (COS) {
                nn0 = fmaf(a00.x, a00.x, fmaf(a00.y, a00.y, fmaf(a01.x, a01.x, fmaf(a01.y, a01.y, nn0))));
                nn1 = fmaf(a10.x, a10.x, fmaf(a10.y, a10.y, fmaf(a11.x, a11.x, fmaf(a11.y, a11.y, nn1))));
            }
            split_pair(a00, ah[0], al[0]);
            split_pair(a10, ah[1], al[1]);
            split_pair(a01, ah[2], al[2]);
            split_pair(a11, ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const size_t off = static_cast<size_t>(nt * 8 + g) * wp + ks * 16 + 2 * t;
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(whi + off), bh1 = *reinterpret_cast<const uint32_t*>(whi + off + 8);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(wlo + off), bl1 = *reinterpret_cast<const uint32_t*>(wlo + off + 8);
                mma_bf16_16816_acc(acc[nt], ah, bh0, bh1);
                mma_bf16_16816_acc(acc[nt], al, bh0, bh1);
                mma_bf16_16816_acc(acc[nt], ah, bl0, bl1);
            }
        }
        if (COS) {
            nn0 += __shfl_xor_sync(0xffffffffu, nn0, 1); nn0 += __shfl_xor_sync(0xffffffffu, nn0, 2);
            nn1 += __shfl_xor_sync(0xffffffffu, nn1, 1); nn1 += __shfl_xor_sync(0xffffffffu, nn1, 2);
            nn0 = fmaxf(sqrtf(nn0), 1e-12f);
            nn1 = fmaxf(sqrtf(nn1), 1e-12f);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = nt * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int cc = c + (e & 1), pp = (e & 2) ? p1 : p0;
                if (cc < C && pp < P) {
                    if (COS) raw[cc * P + pp] = acc[nt][e] / (((e & 2) ? nn1 : nn0) * wnorm[cc]);
                    else raw[cc * P + pp] = relu ? fmaxf(acc[nt][e], 0.f) : acc[nt][e];
                }
            }
        }
    }
    __syncthreads();
    if (COS) {
        float* dst = cam + static_cast<size_t>(b) * C * P;
        for (int i = threadIdx.x; i < C * P; i += blockDim.x) dst[i] = raw[i];
        return;
    }
    for (int c = warp; c < C; c += nwarps) {
        float mn = INFINITY, mx = -INFINITY;
        for (int p = lane; p < P; p += 32) { const float v = raw[c * P + p]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        mn = warp_min(mn);
        mx = warp_max(mx);
        const float den = (mx - mn) + eps;
        float* dst = cam + (static_cast<size_t>(b) * C + c) * P;
        for (int p = lane; p < P; p += 32) dst[p] = (raw[c * P + p] - mn) / den;
    }
}

template <int NT, bool COS = false>
static int launch_cam_project(const float* tokens, const float* w, size_t w_stride, float* cam, int batch, int n_tokens, int dim, int classes, int relu,
                              float eps, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(2) * 8 * NT * (dim + 8) * 2 + sizeof(float) * static_cast<size_t>(classes) * (n_tokens - 1);
    VTC_REQUIRE(smem <= 220 * 1024, VTC_ERR_SHAPE, "cam_project: %zu bytes of smem", smem);
    static SmemOptIn optin;
    int rc_ = optin.ensure(reinterpret_cast<const void*>(cam_project_kernel<NT, COS>), smem);
    if (rc_ != VTC_OK) return rc_;
    int warps = cdiv(n_tokens - 1, 16);
    if (warps > 16) warps = 16;
    if (warps < 1) warps = 1;
    cam_project_kernel<NT, COS><<<batch, warps * 32, smem, stream>>>(tokens, w, w_stride, cam, n_tokens, dim, classes, relu, eps);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

int cam_project(const float* tokens, const float* w, float* cam, int batch, int n_tokens, int dim, int classes, int relu, float eps, cudaStream_t stream) {
    VTC_REQUIRE(tokens && w && cam, VTC_ERR_ARG, "cam_project: null pointer");
    VTC_REQUIRE(batch > 0 && n_tokens > 1 && dim % 128 == 0 && classes > 0 && classes <= 64, VTC_ERR_SHAPE,
                "cam_project: dim %d (multiple of 128) classes %d (<= 64)", dim, classes);
    VTC_REQUIRE(((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(w)) & 7) == 0, VTC_ERR_ARG, "cam_project: pointers must be 8-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    switch (cdiv(classes, 8)) {
        case 1: return launch_cam_project<1>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
        case 2: return launch_cam_project<2>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
        case 3: return launch_cam_project<3>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
        case 4: return launch_cam_project<4>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
        case 5: return launch_cam_project<5>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
        case 6: return launch_cam_project<6>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
        case 7: return launch_cam_project<7>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
        default: return launch_cam_project<8>(tokens, w, 0, cam, batch, n_tokens, dim, classes, relu, eps, stream);
    }
}

// ---- row-wise / max -------------------------------------------------------------------------------------------------
__global__ void normalize_max_kernel(float* __restrict__ maps, int rows, int P) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* m = maps + static_cast<size_t>(row) * P;
    float mx = -INFINITY;
    for (int j = lane; j < P; j += 32) mx = fmaxf(mx, m[j]);
    mx = warp_max(mx);
    for (int j = lane; j < P; j += 32) m[j] = m[j] / mx;
}

int normalize_max(float* maps, int rows, int p, cudaStream_t stream) {
    VTC_REQUIRE(maps, VTC_ERR_ARG, "normalize_max: null pointer");
    VTC_REQUIRE(rows > 0 && p > 0, VTC_ERR_SHAPE, "normalize_max: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    normalize_max_kernel<<<cdiv(rows, 8), 256, 0, stream>>>(maps, rows, p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- bilinear upsampling, align_corners=False (F.interpolate validate.py:177,239 == cv2.resize INTER_LINEAR) ----------
// The three pixel kernels below (full-resolution maps, CAM pseudo label, high-weight-patch segmentation) share one tiling:
// a block owns PIX_ROWS output rows of one image and walks them in quads of four consecutive pixels, so that a thread's result
// leaves as ONE 16-byte (fp32) or 4-byte (uint8) store and a warp writes 512 / 128 contiguous bytes.  Everything that depends
// only on x (source column pair and weight) is tabulated once per block in shared memory next to the g x g source maps; what
// depends only on y is computed once per quad.  The interpolation itself keeps the operation order of F.interpolate /
// cv2.resize (horizontal pair first, then vertical).  Per output pixel and map: 4 shared loads + 6 flops (the first version
// spent > 100 instructions per pixel on 64-bit index arithmetic and reloaded every class map for four pixels per thread).
struct Lerp { int i0, i1; float w0, w1; };
__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in_size) {
    float src = (dst + 0.5f) * scale - 0.5f;
    if (src < 0.f) src = 0.f;
    int i0 = static_cast<int>(src);
    if (i0 > in_size - 1) i0 = in_size - 1;
    const int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    const float l1 = src - i0;
    return Lerp{i0, i1, 1.0f - l1, l1};
}
__device__ __forceinline__ float bilerp(const float* m, int g, const Lerp& y, const Lerp& x) {
    return y.w0 * (x.w0 * m[y.i0 * g + x.i0] + x.w1 * m[y.i0 * g + x.i1]) + y.w1 * (x.w0 * m[y.i1 * g + x.i0] + x.w1 * m[y.i1 * g + x.i1]);
}

constexpr int PIX_ROWS = 32;          // output rows per block
constexpr int PIX_RV = 4;             // consecutive rows per thread item (one quad of columns x PIX_RV rows = 16 pixels)
constexpr int PIX_THREADS = 256;
// Shared memory of a block: xi[W] (byte offset of source column i0) | xw[W] (weight of i0 + 1) | yi[PIX_ROWS] (byte offset of
// source row i0) | yw[PIX_ROWS] (weight of row i0 + 1) | maps.  A source map is staged PADDED, (g+1) x (g+1) floats with its
// last column and row duplicated, so that the second tap is always "+1 column" / "+1 row" (lerp_coord clamps i1 to i0 on the
// border, where both taps then read the same value: identical arithmetic) and a sample needs ONE address computation.
__device__ __forceinline__ int pix_pitch(int g) { return g + 1; }
__device__ __forceinline__ float* pix_fill_tables(uint8_t* sm, int W, int H, int g, int y_first, uint32_t*& xi, float*& xw, uint32_t*& yi, float*& yw) {
    xi = reinterpret_cast<uint32_t*>(sm);
    xw = reinterpret_cast<float*>(xi + W);
    yi = reinterpret_cast<uint32_t*>(xw + W);
    yw = reinterpret_cast<float*>(yi + PIX_ROWS);
    const float sx = static_cast<float>(g) / W, sy = static_cast<float>(g) / H;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const Lerp l = lerp_coord(x, sx, g);
        xi[x] = static_cast<uint32_t>(l.i0) * 4u;
        xw[x] = l.w1;
    }
    if (threadIdx.x < PIX_ROWS) {
        const Lerp l = lerp_coord(min(y_first + static_cast<int>(threadIdx.x), H - 1), sy, g);
        yi[threadIdx.x] = static_cast<uint32_t>(l.i0 * pix_pitch(g)) * 4u;
        yw[threadIdx.x] = l.w1;
    }
    return yw + PIX_ROWS;
}
// stage `n` g x g maps (map k at src + k * src_stride floats) padded into dst
__device__ __forceinline__ void pix_stage_maps(float* dst, const float* __restrict__ src, size_t src_stride, int n, int g) {
    const int pitch = pix_pitch(g), pp = pitch * pitch;
    for (int i = threadIdx.x; i < n * pp; i += blockDim.x) {
        const int k = i / pp, r = i - k * pp, y = r / pitch, x = r - y * pitch;
        dst[i] = src[k * src_stride + min(y, g - 1) * g + min(x, g - 1)];
    }
}
static size_t pix_table_bytes(int W) { return static_cast<size_t>(W) * 8 + PIX_ROWS * 8; }
static size_t pix_map_bytes(int g, int n) { return sizeof(float) * static_cast<size_t>(n) * (g + 1) * (g + 1); }

// One thread item: rows r0 .. r0+3 of the block x the four columns of one quad.
struct PixItem {
    uint32_t xo[4];        // byte offsets of the source columns
    float xw[4];           // weights of column i0 + 1
    uint32_t yo[PIX_RV];   // byte offsets of the source rows
    float yw[PIX_RV];      // weights of row i0 + 1
    bool same;             // the four rows lie in the same source cell
    int r0, x0;
};
// item index -> item; false when it starts below the last row of the tile
__device__ __forceinline__ bool pix_decode(int it, int quads, int rows, int W, const uint32_t* xi, const float* xw, const uint32_t* yi, const float* yw, PixItem& p) {
    const int rg = it / quads;
    p.r0 = rg * PIX_RV;
    if (p.r0 >= rows) return false;
    p.x0 = (it - rg * quads) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int x = min(p.x0 + e, W - 1);
        p.xo[e] = xi[x];
        p.xw[e] = xw[x];
    }
#pragma unroll
    for (int r = 0; r < PIX_RV; ++r) {
        p.yo[r] = yi[p.r0 + r];
        p.yw[r] = yw[p.r0 + r];
    }
    p.same = (p.yo[1] == p.yo[0]) && (p.yo[2] == p.yo[0]) && (p.yo[3] == p.yo[0]);
    return true;
}
// The 16 bilinear samples of an item of ONE padded map m: f(r, e, value).  When the four rows lie in the same source cell (26
// of 27 row groups at 375 rows from a 14 x 14 map) the horizontal interpolation of the two source rows is done once for the
// item (16 instead of 64 loads); the operation order per sample is the reference's in both paths (horizontal pair first, then
// vertical).
template <typename Fn>
__device__ __forceinline__ void pix_item(const float* m, uint32_t pitch_bytes, const PixItem& p, Fn&& f) {
    const uint8_t* mb = reinterpret_cast<const uint8_t*>(m);
    if (p.same) {
        float ha[4], hb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float* a = reinterpret_cast<const float*>(mb + p.yo[0] + p.xo[e]);
            const float* c = reinterpret_cast<const float*>(mb + p.yo[0] + p.xo[e] + pitch_bytes);
            const float w1 = p.xw[e], w0 = 1.0f - w1;
            ha[e] = w0 * a[0] + w1 * a[1];
            hb[e] = w0 * c[0] + w1 * c[1];
        }
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) {
            const float v1 = p.yw[r], v0 = 1.0f - v1;
#pragma unroll
            for (int e = 0; e < 4; ++e) f(r, e, v0 * ha[e] + v1 * hb[e]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) {
            const float v1 = p.yw[r], v0 = 1.0f - v1;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float* a = reinterpret_cast<const float*>(mb + p.yo[r] + p.xo[e]);
                const float* c = reinterpret_cast<const float*>(mb + p.yo[r] + p.xo[e] + pitch_bytes);
                const float w1 = p.xw[e], w0 = 1.0f - w1;
                f(r, e, v0 * (w0 * a[0] + w1 * a[1]) + v1 * (w0 * c[0] + w1 * c[1]));
            }
        }
    }
}
__device__ __forceinline__ void pix_store_u8(uint8_t* dst, uint32_t packed, int x0, int W, bool vec) {
    if (vec) *reinterpret_cast<uint32_t*>(dst) = packed;
    else for (int e = 0; e < 4 && x0 + e < W; ++e) dst[e] = static_cast<uint8_t>(packed >> (8 * e));
}

template <bool U8>
__global__ void __launch_bounds__(PIX_THREADS) upsample_kernel(const float* __restrict__ in, void* __restrict__ out, int g, int H, int W) {
    extern __shared__ __align__(16) uint8_t pix_sm[];
    uint32_t *xi, *yi;
    float *xw, *yw;
    const int y_first = blockIdx.x * PIX_ROWS, rows = min(PIX_ROWS, H - y_first);
    float* m = pix_fill_tables(pix_sm, W, H, g, y_first, xi, xw, yi, yw);      // one padded map
    const int n = blockIdx.y;
    pix_stage_maps(m, in + static_cast<size_t>(n) * g * g, 0, 1, g);
    __syncthreads();
    const uint32_t pitch_bytes = pix_pitch(g) * 4;
    const int quads = (W + 3) >> 2;
    const size_t base = static_cast<size_t>(n) * H * W;
    const bool vec = (W & 3) == 0;
    for (int it = threadIdx.x; it < (PIX_ROWS / PIX_RV) * quads; it += blockDim.x) {
        PixItem p;
        if (!pix_decode(it, quads, rows, W, xi, xw, yi, yw, p)) break;
        const int r0 = p.r0, x0 = p.x0;
        float v[PIX_RV][4];
        pix_item(m, pitch_bytes, p, [&](int r, int e, float s) { v[r][e] = s; });
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) {
            if (r0 + r >= rows) break;
            const size_t o = base + static_cast<size_t>(y_first + r0 + r) * W + x0;
            if (U8) {                                      // .astype("uint8") truncation, predict.py:269
                const uint32_t p = static_cast<uint32_t>(static_cast<uint8_t>(v[r][0] * 255.0f)) | (static_cast<uint32_t>(static_cast<uint8_t>(v[r][1] * 255.0f)) << 8) |
                                   (static_cast<uint32_t>(static_cast<uint8_t>(v[r][2] * 255.0f)) << 16) | (static_cast<uint32_t>(static_cast<uint8_t>(v[r][3] * 255.0f)) << 24);
                pix_store_u8(static_cast<uint8_t*>(out) + o, p, x0, W, vec);
            } else {
                float* dst = static_cast<float*>(out) + o;
                if (vec) st_f4(dst, make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
                else for (int e = 0; e < 4 && x0 + e < W; ++e) dst[e] = v[r][e];
            }
        }
    }
}

template <bool U8>
static int upsample(const float* in, void* out, int n, int g, int H, int W, cudaStream_t stream) {
    VTC_REQUIRE(in && out, VTC_ERR_ARG, "upsample: null pointer");
    const size_t smem = pix_table_bytes(W) + pix_map_bytes(g, 1);
    VTC_REQUIRE(n > 0 && g > 0 && g < 1024 && H > 0 && W > 0 && smem <= 48 * 1024 && n <= 65535, VTC_ERR_SHAPE, "upsample: bad shape n=%d g=%d W=%d", n, g, W);
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, VTC_ERR_ARG, "upsample: output must be 16-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    upsample_kernel<U8><<<dim3(cdiv(H, PIX_ROWS), n), PIX_THREADS, smem, stream>>>(in, out, g, H, W);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- CAM pseudo label: fused upsample + argmax over [bg_thresh, labelled classes] -----------------------------------
// Only the maps of the image's own classes (utils.py:100-108; 1.5 per VOC image on average) are staged in shared memory.
__global__ void __launch_bounds__(PIX_THREADS) cam_label_kernel(const float* __restrict__ cam, const uint8_t* __restrict__ labels, float bg_thresh,
                                                               uint8_t* __restrict__ out, int C, int g, int H, int W) {
    extern __shared__ __align__(16) uint8_t pix_sm[];
    __shared__ int active[64];
    __shared__ int nactive;
    uint32_t *xi, *yi;
    float *xw, *yw;
    const int y_first = blockIdx.x * PIX_ROWS, rows = min(PIX_ROWS, H - y_first);
    float* m = pix_fill_tables(pix_sm, W, H, g, y_first, xi, xw, yi, yw);      // [nactive] padded maps
    const int b = blockIdx.y;
    const int gg = g * g, pp = pix_pitch(g) * pix_pitch(g);
    const uint32_t pitch_bytes = pix_pitch(g) * 4;
    if (threadIdx.x == 0) {
        int k = 0;
        for (int c = 0; c < C; ++c)
            if (labels[b * C + c]) active[k++] = c;
        nactive = k;
    }
    __syncthreads();
    const int na = nactive;
    for (int k = 0; k < na; ++k) pix_stage_maps(m + k * pp, cam + (static_cast<size_t>(b) * C + active[k]) * gg, 0, 1, g);
    __syncthreads();
    const int quads = (W + 3) >> 2;
    const size_t base = static_cast<size_t>(b) * H * W;
    const bool vec = (W & 3) == 0;
    for (int it = threadIdx.x; it < (PIX_ROWS / PIX_RV) * quads; it += blockDim.x) {
        PixItem p;
        if (!pix_decode(it, quads, rows, W, xi, xw, yi, yw, p)) break;
        const int r0 = p.r0, x0 = p.x0;
        float best[PIX_RV][4];
        uint32_t lab[PIX_RV];                              // four labels per row, one per byte
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) {
            lab[r] = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) best[r][e] = bg_thresh;
        }
        for (int k = 0; k < na; ++k) {
            const uint32_t code = static_cast<uint32_t>(active[k] + 1);
            pix_item(m + k * pp, pitch_bytes, p, [&](int r, int e, float s) {
                if (s > best[r][e]) {                      // strict: ties keep the earlier entry like torch.argmax
                    best[r][e] = s;
                    lab[r] = (lab[r] & ~(0xffu << (8 * e))) | (code << (8 * e));
                }
            });
        }
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) {
            if (r0 + r >= rows) break;
            pix_store_u8(out + base + static_cast<size_t>(y_first + r0 + r) * W + x0, lab[r], x0, W, vec);
        }
    }
}

int cam_label(const float* cam, const uint8_t* labels, float bg_thresh, uint8_t* out, int batch, int classes, int g, int H, int W, cudaStream_t stream) {
    VTC_REQUIRE(cam && labels && out, VTC_ERR_ARG, "cam_label: null pointer");
    VTC_REQUIRE(batch > 0 && batch <= 65535 && classes > 0 && classes <= 64 && g > 0 && g < 1024 && H > 0 && W > 0, VTC_ERR_SHAPE, "cam_label: bad shape");
    const size_t smem = pix_table_bytes(W) + pix_map_bytes(g, classes);
    VTC_REQUIRE(smem <= 48 * 1024, VTC_ERR_SHAPE, "cam_label: %zu bytes of smem", smem);
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 3) == 0, VTC_ERR_ARG, "cam_label: output must be 4-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    cam_label_kernel<<<dim3(cdiv(H, PIX_ROWS), batch), PIX_THREADS, smem, stream>>>(cam, labels, bg_thresh, out, classes, g, H, W);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- high-weight-patch class vote + cosine maps (validate.py:132-175) -------------------------------------------------
constexpr int HWP_MAXK = 16;
// One block per image: class vote of the K high-weight patches (validate.py:132-153).  The cosine maps (validate.py:157-175)
// are a [P x D] x [D x K] product per image with per-image weights: cam_project_kernel<.., COS> above (warp-level tensor cores,
// fp32-equivalent split operands), which replaced a shared-memory-bound FMA version of this kernel (282 -> see profiles/).
__global__ void __launch_bounds__(256) hwp_vote_kernel(const float* __restrict__ hwp_logits, const float* __restrict__ w1, const float* __restrict__ ori,
                                                       float sig_thresh, int32_t* __restrict__ p2c, int D, int C, int K) {
    __shared__ int votes[HWP_MAXK * 64];   // [K][C]
    __shared__ int pred[64];
    const int b = blockIdx.x;
    const float* ob = ori + static_cast<size_t>(b) * K * D;
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) votes[i] = 0;
    if (threadIdx.x < C) pred[threadIdx.x] = (1.0f / (1.0f + expf(-hwp_logits[b * C + threadIdx.x]))) >= sig_thresh;   // validate.py:132-134
    __syncthreads();
    // feature vote: class of feature d (argmax over predicted classes' head1 rows) goes to the hw patch owning d
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float bw = -INFINITY;
        int bc = 0;
        for (int c = 0; c < C; ++c) {
            const float v = pred[c] ? __ldg(w1 + static_cast<size_t>(c) * D + d) : -10.0f;     // validate.py:138-143
            if (v > bw) { bw = v; bc = c; }
        }
        float bo = -INFINITY;
        int bk = 0;
        for (int k = 0; k < K; ++k) {
            const float v = __ldg(ob + static_cast<size_t>(k) * D + d);
            if (v > bo) { bo = v; bk = k; }                                       // validate.py:148
        }
        atomicAdd(&votes[bk * C + bc], 1);
    }
    __syncthreads();
    if (threadIdx.x < K) {
        int best = 0, bc = -1;
        for (int c = 0; c < C; ++c)
            if (votes[threadIdx.x * C + c] > best) { best = votes[threadIdx.x * C + c]; bc = c; }   // mode, ties -> smallest class
        p2c[b * K + threadIdx.x] = bc;                                            // -1: patch owns no feature
    }
}

int hwp_cos_vote(const float* hwp_logits, const float* head1_w, const float* hwp_tokens, const float* tokens, float sig_thresh,
                 int32_t* patch_to_cls, float* cosm, int batch, int n_tokens, int dim, int classes, int k, cudaStream_t stream) {
    VTC_REQUIRE(hwp_logits && head1_w && hwp_tokens && tokens && patch_to_cls && cosm, VTC_ERR_ARG, "hwp_cos_vote: null pointer");
    VTC_REQUIRE(batch > 0 && n_tokens > 1 && dim % 128 == 0 && classes > 0 && classes <= 64 && k > 0 && k <= HWP_MAXK, VTC_ERR_SHAPE,
                "hwp_cos_vote: dim %d classes %d k %d", dim, classes, k);
    VTC_REQUIRE(((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(hwp_tokens)) & 7) == 0, VTC_ERR_ARG, "hwp_cos_vote: pointers must be 8-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    hwp_vote_kernel<<<batch, 256, 0, stream>>>(hwp_logits, head1_w, hwp_tokens, sig_thresh, patch_to_cls, dim, classes, k);
    VTC_CHECK_LAUNCH();
    const size_t stride = static_cast<size_t>(k) * dim;
    if (k <= 8) return launch_cam_project<1, true>(tokens, hwp_tokens, stride, cosm, batch, n_tokens, dim, k, 0, 0.f, stream);
    return launch_cam_project<2, true>(tokens, hwp_tokens, stride, cosm, batch, n_tokens, dim, k, 0, 0.f, stream);
}

// ---- validate.py:177-258 fused: upsample K cosine maps + argmax + fg/bg thresholds -> uint8 label map ------------------
// Same tiling as above; K + 1 maps per image in shared memory, K bilinear samples per pixel (inherent: the argmax is taken at
// output resolution, validate.py:177-180).
__global__ void __launch_bounds__(PIX_THREADS) hwp_seg_kernel(const float* __restrict__ cosm, const int32_t* __restrict__ p2c, const float* __restrict__ bg_map,
                                                             float cos_thresh, float bg_thresh, uint8_t* __restrict__ out, int K, int g, int H, int W) {
    extern __shared__ __align__(16) uint8_t pix_sm[];
    __shared__ int cls_s[HWP_MAXK];
    uint32_t *xi, *yi;
    float *xw, *yw;
    const int y_first = blockIdx.x * PIX_ROWS, rows = min(PIX_ROWS, H - y_first);
    float* m = pix_fill_tables(pix_sm, W, H, g, y_first, xi, xw, yi, yw);      // K padded cosine maps + the padded background map
    const int b = blockIdx.y;
    const int gg = g * g, pp = pix_pitch(g) * pix_pitch(g);
    const uint32_t pitch_bytes = pix_pitch(g) * 4;
    pix_stage_maps(m, cosm + static_cast<size_t>(b) * K * gg, gg, K, g);
    pix_stage_maps(m + K * pp, bg_map + static_cast<size_t>(b) * gg, 0, 1, g);
    if (threadIdx.x < K) cls_s[threadIdx.x] = p2c[b * K + threadIdx.x];
    __syncthreads();
    const int quads = (W + 3) >> 2;
    const size_t base = static_cast<size_t>(b) * H * W;
    const bool vec = (W & 3) == 0;
    for (int it = threadIdx.x; it < (PIX_ROWS / PIX_RV) * quads; it += blockDim.x) {
        PixItem p;
        if (!pix_decode(it, quads, rows, W, xi, xw, yi, yw, p)) break;
        const int r0 = p.r0, x0 = p.x0;
        float best[PIX_RV][4];
        uint32_t bk[PIX_RV];                               // arg max per pixel, one byte each
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) {
            bk[r] = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) best[r][e] = -INFINITY;
        }
        for (int k = 0; k < K; ++k) {
            pix_item(m + k * pp, pitch_bytes, p, [&](int r, int e, float s) {
                if (s > best[r][e]) {                      // validate.py:179-180
                    best[r][e] = s;
                    bk[r] = (bk[r] & ~(0xffu << (8 * e))) | (static_cast<uint32_t>(k) << (8 * e));
                }
            });
        }
        uint32_t lab[PIX_RV];
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) lab[r] = 0u;
        pix_item(m + K * pp, pitch_bytes, p, [&](int r, int e, float s) {
            const bool fg = best[r][e] >= cos_thresh;                              // validate.py:183-186
            const bool keep = s >= bg_thresh;                                      // validate.py:239-246
            const int c = cls_s[(bk[r] >> (8 * e)) & 0xffu];
            lab[r] |= ((fg && keep && c >= 0) ? static_cast<uint32_t>(c + 1) : 0u) << (8 * e);    // validate.py:190-258
        });
#pragma unroll
        for (int r = 0; r < PIX_RV; ++r) {
            if (r0 + r >= rows) break;
            pix_store_u8(out + base + static_cast<size_t>(y_first + r0 + r) * W + x0, lab[r], x0, W, vec);
        }
    }
}

int hwp_seg(const float* cosm, const int32_t* p2c, const float* bg_map, float cos_thresh, float bg_thresh, uint8_t* out, int batch, int k, int g,
            int H, int W, cudaStream_t stream) {
    VTC_REQUIRE(cosm && p2c && bg_map && out, VTC_ERR_ARG, "hwp_seg: null pointer");
    VTC_REQUIRE(batch > 0 && batch <= 65535 && k > 0 && k <= HWP_MAXK && g > 0 && g < 1024 && H > 0 && W > 0, VTC_ERR_SHAPE, "hwp_seg: bad shape");
    const size_t smem = pix_table_bytes(W) + pix_map_bytes(g, k + 1);
    VTC_REQUIRE(smem <= 48 * 1024, VTC_ERR_SHAPE, "hwp_seg: %zu bytes of smem", smem);
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 3) == 0, VTC_ERR_ARG, "hwp_seg: output must be 4-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    hwp_seg_kernel<<<dim3(cdiv(H, PIX_ROWS), batch), PIX_THREADS, smem, stream>>>(cosm, p2c, bg_map, cos_thresh, bg_thresh, out, k, g, H, W);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- confusion matrix (utils.py:35-45) -------------------------------------------------------------------------------------
__global__ void confmat_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred, size_t count, int n, unsigned long long* __restrict__ mat) {
    extern __shared__ unsigned int hist[];    // [n*n]
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
        const int a = gt[i], p = pred[i];
        if (a < n && p < n) atomicAdd(&hist[a * n + p], 1u);      // k = (a >= 0) & (a < n): 255 = ignore
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += blockDim.x)
        if (hist[i]) atomicAdd(&mat[i], static_cast<unsigned long long>(hist[i]));
}

int confmat_update(const uint8_t* gt, const uint8_t* pred, size_t count, int n, int64_t* mat, cudaStream_t stream) {
    VTC_REQUIRE(gt && pred && mat, VTC_ERR_ARG, "confmat: null pointer");
    VTC_REQUIRE(n > 0 && n <= 64, VTC_ERR_SHAPE, "confmat: n=%d", n);
    if (count == 0) return VTC_OK;
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int grid = grid_cap((count + 256 * 16 - 1) / (256 * 16), 8);
    confmat_kernel<<<grid, 256, sizeof(unsigned int) * n * n, stream>>>(gt, pred, count, n, reinterpret_cast<unsigned long long*>(mat));
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- per-image average precision (utils.py:248-262 -> sklearn.metrics.average_precision_score) --------------------------
// AP = sum over the distinct score thresholds t (descending) of (R(t) - R(prev)) * P(t), with P = tp / #(score >= t) and
// R = tp / #positives.  C is tiny (20): one thread per image, O(C^2), fp64 like the numpy path.  Images without a positive
// label are skipped (utils.py:256); acc[0] += AP, acc[1] += 1 for the others.
__global__ void average_precision_kernel(const float* __restrict__ labels, const float* __restrict__ scores, int B, int C,
                                         double* __restrict__ ap, double* __restrict__ acc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* y = labels + static_cast<size_t>(b) * C;
    const float* sc = scores + static_cast<size_t>(b) * C;
    double npos = 0.0;
    for (int j = 0; j < C; ++j) npos += (y[j] != 0.f) ? 1.0 : 0.0;
    if (npos == 0.0) {
        if (ap) ap[b] = -1.0;
        return;
    }
    double total = 0.0;
    for (int i = 0; i < C; ++i) {
        const float t = sc[i];
        bool first = true;                      // one term per distinct threshold
        for (int j = 0; j < i; ++j) first = first && (sc[j] != t);
        if (!first) continue;
        double tp = 0.0, tp_above = 0.0, cnt = 0.0;
        for (int j = 0; j < C; ++j) {
            const double pos = (y[j] != 0.f) ? 1.0 : 0.0;
            if (sc[j] >= t) { tp += pos; cnt += 1.0; }
            if (sc[j] > t) tp_above += pos;
        }
        total += ((tp - tp_above) / npos) * (tp / cnt);
    }
    if (ap) ap[b] = total;
    if (acc) {
        atomicAdd(&acc[0], total);
        atomicAdd(&acc[1], 1.0);
    }
}

int average_precision(const float* labels, const float* scores, int batch, int classes, double* ap, double* acc, cudaStream_t stream) {
    VTC_REQUIRE(labels && scores && (ap || acc), VTC_ERR_ARG, "average_precision: null pointer");
    VTC_REQUIRE(batch > 0 && classes > 0, VTC_ERR_SHAPE, "average_precision: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    average_precision_kernel<<<cdiv(batch, 128), 128, 0, stream>>>(labels, scores, batch, classes, ap, acc);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- patch-token similarity matrix of predict.py:191-199 ------------------------------------------------------------------
// The reference calls F.normalize on a [1,N,D] tensor with the default dim=1, i.e. every FEATURE column is L2-normalised
// across the N tokens (not every token across its features), then takes the N x N gram matrix.  Reproduced as is
// (faithful quirk): sim[b,i,j] = sum_d x[b,i,d] x[b,j,d] / max(||x[b,:,d]||, 1e-12)^2.
__global__ void feature_norm_kernel(const float* __restrict__ x, float* __restrict__ inv_sq, int N, int D) {
    const int b = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const float* src = x + static_cast<size_t>(b) * N * D + d;
    float s = 0.f;
    for (int n = 0; n < N; ++n) { const float v = src[static_cast<size_t>(n) * D]; s = fmaf(v, v, s); }
    const float nrm = fmaxf(sqrtf(s), 1e-12f);
    inv_sq[static_cast<size_t>(b) * D + d] = 1.0f / (nrm * nrm);
}
// grid (N, B), block 256: row i of image b against every row j; x_i * inv_sq staged in smem
__global__ void patch_similarity_kernel(const float* __restrict__ x, const float* __restrict__ inv_sq, float* __restrict__ sim, int N, int D) {
    extern __shared__ float xi[];     // [D]
    const int i = blockIdx.x, b = blockIdx.y;
    const float* xb = x + static_cast<size_t>(b) * N * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) xi[d] = xb[static_cast<size_t>(i) * D + d] * inv_sq[static_cast<size_t>(b) * D + d];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < N; j += nwarps) {
        const float* xj = xb + static_cast<size_t>(j) * D;
        float s = 0.f;
        for (int d = lane * 4; d < D; d += 128) {
            const float4 v = ldg_f4(xj + d);
            s = fmaf(v.x, xi[d], fmaf(v.y, xi[d + 1], fmaf(v.z, xi[d + 2], fmaf(v.w, xi[d + 3], s))));
        }
        s = warp_sum(s);
        if (lane == 0) sim[(static_cast<size_t>(b) * N + i) * N + j] = s;
    }
}

int patch_similarity(const float* tokens, float* scratch, float* sim, int batch, int n_tokens, int dim, cudaStream_t stream) {
    VTC_REQUIRE(tokens && scratch && sim, VTC_ERR_ARG, "patch_similarity: null pointer");
    VTC_REQUIRE(batch > 0 && batch <= 65535 && n_tokens > 0 && dim > 0 && dim % 128 == 0 && dim * 4 <= 48 * 1024, VTC_ERR_SHAPE, "patch_similarity: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    feature_norm_kernel<<<dim3(cdiv(dim, 128), batch), 128, 0, stream>>>(tokens, scratch, n_tokens, dim);
    VTC_CHECK_LAUNCH();
    patch_similarity_kernel<<<dim3(n_tokens, batch), 256, sizeof(float) * dim, stream>>>(tokens, scratch, sim, n_tokens, dim);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

}  // namespace vtc

extern "C" {
int vtc_average_precision(const float* labels, const float* scores, int32_t batch, int32_t classes, double* ap, double* acc, void* stream) {
    return vtc::average_precision(labels, scores, batch, classes, ap, acc, static_cast<cudaStream_t>(stream));
}
int vtc_patch_similarity(const float* tokens, float* scratch, float* sim, int32_t batch, int32_t n_tokens, int32_t dim, void* stream) {
    return vtc::patch_similarity(tokens, scratch, sim, batch, n_tokens, dim, static_cast<cudaStream_t>(stream));
}
int vtc_rollout(const float* attn_mean, float* row, int32_t layers, int32_t batch, int32_t n_tokens, void* stream) {
    return vtc::rollout(attn_mean, row, layers, batch, n_tokens, static_cast<cudaStream_t>(stream));
}
int32_t vtc_rollout_operand_ld(int32_t n_tokens) { return vtc::rollout_operand_ld(n_tokens); }
int vtc_rollout_operand_from_mean(const float* attn_mean, void* operand, int32_t batch, int32_t n_tokens, void* stream) {
    return vtc::rollout_operand_from_mean(attn_mean, operand, batch, n_tokens, static_cast<cudaStream_t>(stream));
}
int vtc_rollout_operands(const void* operands, float* row, int32_t layers, int32_t batch, int32_t n_tokens, void* stream) {
    return vtc::rollout_operand(operands, row, layers, batch, n_tokens, static_cast<cudaStream_t>(stream));
}
int vtc_cls_layer_map(const float* cls_rows, float* map, int32_t layers, int32_t first, int32_t last, int32_t batch, int32_t heads,
                      int32_t n_tokens, void* stream) {
    return vtc::cls_layer_map(cls_rows, map, layers, first, last, batch, heads, n_tokens, static_cast<cudaStream_t>(stream));
}
int vtc_cam_project(const float* tokens, const float* w, float* cam, int32_t batch, int32_t n_tokens, int32_t dim, int32_t classes,
                    int32_t relu, float eps, void* stream) {
    return vtc::cam_project(tokens, w, cam, batch, n_tokens, dim, classes, relu, eps, static_cast<cudaStream_t>(stream));
}
int vtc_normalize_max(float* maps, int32_t rows, int32_t p, void* stream) {
    return vtc::normalize_max(maps, rows, p, static_cast<cudaStream_t>(stream));
}
int vtc_upsample_bilinear(const float* in, float* out, int32_t n, int32_t g, int32_t out_h, int32_t out_w, void* stream) {
    return vtc::upsample<false>(in, out, n, g, out_h, out_w, static_cast<cudaStream_t>(stream));
}
int vtc_upsample_bilinear_u8(const float* in, uint8_t* out, int32_t n, int32_t g, int32_t out_h, int32_t out_w, void* stream) {
    return vtc::upsample<true>(in, out, n, g, out_h, out_w, static_cast<cudaStream_t>(stream));
}
int vtc_cam_label(const float* cam, const uint8_t* labels, float bg_thresh, uint8_t* out, int32_t batch, int32_t classes, int32_t g,
                  int32_t out_h, int32_t out_w, void* stream) {
    return vtc::cam_label(cam, labels, bg_thresh, out, batch, classes, g, out_h, out_w, static_cast<cudaStream_t>(stream));
}
int vtc_hwp_cos_vote(const float* hwp_logits, const float* head1_w, const float* hwp_tokens, const float* tokens, float sig_thresh,
                     int32_t* patch_to_cls, float* cos, int32_t batch, int32_t n_tokens, int32_t dim, int32_t classes, int32_t k, void* stream) {
    return vtc::hwp_cos_vote(hwp_logits, head1_w, hwp_tokens, tokens, sig_thresh, patch_to_cls, cos, batch, n_tokens, dim, classes, k,
                             static_cast<cudaStream_t>(stream));
}
int vtc_hwp_seg(const float* cos, const int32_t* patch_to_cls, const float* bg_map, float cos_thresh, float bg_thresh, uint8_t* out,
                int32_t batch, int32_t k, int32_t g, int32_t out_h, int32_t out_w, void* stream) {
    return vtc::hwp_seg(cos, patch_to_cls, bg_map, cos_thresh, bg_thresh, out, batch, k, g, out_h, out_w, static_cast<cudaStream_t>(stream));
}
int vtc_confmat_update(const uint8_t* gt, const uint8_t* pred, size_t count, int32_t n, int64_t* mat, void* stream) {
    return vtc::confmat_update(gt, pred, count, n, mat, static_cast<cudaStream_t>(stream));
}
}
