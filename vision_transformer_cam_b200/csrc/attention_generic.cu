// Attention for head dimensions other than 64 (vit_model.py:113-137 with head_dim = embed_dim / num_heads: ViT-H/14 has
// 1280 / 16 = 80).  The tcgen05 kernels (attention_cs / attention_kv / attention.cu) are built around one 128-byte swizzle
// atom per Q / K / V row, i.e. exactly 64 bf16; this kernel is the general-shape path of the same operator: bf16 operands,
// fp32 products / softmax / P (P is NOT rounded to bf16 before P V here), every output of vtc_attention -- O, the CLS query
// row and, on request, the full P.  It runs on the FMA pipe out of shared memory; ViT-H is not a benchmark configuration
// of the reference (whose forward cannot run it at all: 197 tokens and 12 heads are hard-coded), so the
// point here is the complete factory surface, not speed.
//
// One CTA = 32 query rows of one (image, head); 8 warps x 4 rows.  K ([N][hd+2] bf16: odd word pitch, conflict-free for
// lane = key), V ([N][hd] bf16: lane = feature) and the 32 Q rows ([warp][hd][4] fp32: one broadcast 16-byte read gives a
// feature of all four rows) sit in shared memory.
//   phase 1: lane owns keys lane + 32 t: s[4][T] = q . k (8 FMAs per key pair of features), + mask, softmax across the warp;
//   phase 2: P goes through a per-warp staging area ([key][4] fp32), lane owns features lane + 32 u: o[4][U] += p * v.
#include "common.cuh"
#include "ops.h"

namespace vtc {

namespace ag {
constexpr int QT = 32;            // query rows per CTA
constexpr int WARPS = 8;
constexpr int ROWS = QT / WARPS;  // 4 rows per warp
constexpr int TMAX = 10;          // keys per lane: n_tokens <= 320
constexpr int UMAX = 4;           // features per lane: head_dim <= 128

struct Params {
    const __nv_bfloat16* qkv;   // [B,N,3,H,hd]
    const float* key_bias;      // [B,N] or null
    __nv_bfloat16* out;         // [B,N,H*hd]
    float* cls_rows;            // [B,H,N] or null
    float* attn;                // [B,H,N,N] or null
    int B, N, H, hd;
    int npad;                   // N rounded up to 32
    float scale;
};

__host__ __device__ inline size_t smem_bytes(int npad, int hd) {
    return static_cast<size_t>(npad) * (hd + 2) * 2 + static_cast<size_t>(npad) * hd * 2 + static_cast<size_t>(QT) * hd * 4 +
           static_cast<size_t>(WARPS) * npad * ROWS * 4;
}

__global__ void __launch_bounds__(WARPS * 32) attention_generic_kernel(const Params p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int N = p.N, H = p.H, hd = p.hd, npad = p.npad;
    const int kp = hd + 2;                                       // K row pitch in bf16
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* Vs = Ks + static_cast<size_t>(npad) * kp;
    float* Qs = reinterpret_cast<float*>(Vs + static_cast<size_t>(npad) * hd);      // [WARPS][hd][ROWS]
    float* Ps = Qs + QT * hd;                                                       // [WARPS][npad][ROWS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
    const size_t tok_stride = static_cast<size_t>(3) * H * hd;                      // elements per token in qkv
    const __nv_bfloat16* base = p.qkv + static_cast<size_t>(b) * N * tok_stride + static_cast<size_t>(h) * hd;

    // ---- stage K, V (all keys of this head) and the 32 query rows
    const int pairs = hd >> 1;
    for (int i = threadIdx.x; i < npad * pairs; i += blockDim.x) {
        const int n = i / pairs, d2 = i - n * pairs;
        uint32_t kv = 0u, vv = 0u;
        if (n < N) {
            const __nv_bfloat16* tok = base + static_cast<size_t>(n) * tok_stride;
            kv = *reinterpret_cast<const uint32_t*>(tok + static_cast<size_t>(H) * hd + 2 * d2);
            vv = *reinterpret_cast<const uint32_t*>(tok + static_cast<size_t>(2) * H * hd + 2 * d2);
        }
        *reinterpret_cast<uint32_t*>(Ks + static_cast<size_t>(n) * kp + 2 * d2) = kv;
        *reinterpret_cast<uint32_t*>(Vs + static_cast<size_t>(n) * hd + 2 * d2) = vv;
    }
    for (int i = threadIdx.x; i < QT * hd; i += blockDim.x) {
        const int r = i / hd, d = i - r * hd;                    // r: row inside the tile
        const int row = q0 + r;
        const float v = row < N ? __bfloat162float(base[static_cast<size_t>(row) * tok_stride + d]) : 0.f;
        Qs[(static_cast<size_t>(r / ROWS) * hd + d) * ROWS + (r % ROWS)] = v;
    }
    __syncthreads();

    const int row0 = q0 + warp * ROWS;                           // first of this warp's four rows
    if (row0 >= N) return;
    const int T = npad >> 5;
    const float* kb = p.key_bias ? p.key_bias + static_cast<size_t>(b) * N : nullptr;

    // ---- phase 1: scores of my keys for the four rows
    float s[ROWS][TMAX];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int t = 0; t < TMAX; ++t) s[r][t] = 0.f;
    const float4* q4 = reinterpret_cast<const float4*>(Qs + static_cast<size_t>(warp) * hd * ROWS);
    for (int d2 = 0; d2 < pairs; ++d2) {
        const float4 qa = q4[2 * d2], qb = q4[2 * d2 + 1];
#pragma unroll
        for (int t = 0; t < TMAX; ++t) {
            if (t < T) {
                const uint32_t kk = *reinterpret_cast<const uint32_t*>(Ks + static_cast<size_t>(lane + 32 * t) * kp + 2 * d2);
                const float k0 = __uint_as_float(kk << 16), k1 = __uint_as_float(kk & 0xffff0000u);
                s[0][t] = fmaf(qa.x, k0, fmaf(qb.x, k1, s[0][t]));
                s[1][t] = fmaf(qa.y, k0, fmaf(qb.y, k1, s[1][t]));
                s[2][t] = fmaf(qa.z, k0, fmaf(qb.z, k1, s[2][t]));
                s[3][t] = fmaf(qa.w, k0, fmaf(qb.w, k1, s[3][t]));
            }
        }
    }
    // ---- scale, mask (-100 on background keys for foreground query rows, vit_model.py:348-361; a background row's uniform
    //      -100 is softmax-invariant), softmax
    float* Pw = Ps + static_cast<size_t>(warp) * npad * ROWS;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int row = row0 + r;
        const bool row_ok = row < N;
        const bool row_fg = kb == nullptr || !row_ok || kb[row] == 0.f;
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < TMAX; ++t) {
            const int j = lane + 32 * t;
            if (t < T && j < N) {
                float x = s[r][t] * p.scale;
                if (kb != nullptr && row_fg) x += kb[j];
                s[r][t] = x;
                mx = fmaxf(mx, x);
            } else {
                s[r][t] = -INFINITY;
            }
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < TMAX; ++t) {
            const float e = (t < T && s[r][t] > -INFINITY) ? __expf(s[r][t] - mx) : 0.f;
            s[r][t] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int t = 0; t < TMAX; ++t) {
            if (t < T) {
                const int j = lane + 32 * t;
                const float pv = s[r][t] * inv;
                Pw[static_cast<size_t>(j) * ROWS + r] = pv;
                if (row_ok && j < N) {
                    if (p.attn != nullptr) p.attn[((static_cast<size_t>(b) * H + h) * N + row) * N + j] = pv;
                    if (row == 0 && p.cls_rows != nullptr) p.cls_rows[(static_cast<size_t>(b) * H + h) * N + j] = pv;
                }
            }
        }
    }
    __syncwarp();

    // ---- phase 2: O = P V, lane owns features lane + 32 u
    float o[ROWS][UMAX];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int u = 0; u < UMAX; ++u) o[r][u] = 0.f;
    const float4* p4 = reinterpret_cast<const float4*>(Pw);
    for (int j = 0; j < N; ++j) {
        const float4 pj = p4[j];
#pragma unroll
        for (int u = 0; u < UMAX; ++u) {
            const int d = lane + 32 * u;
            if (d < hd) {
                const float v = __bfloat162float(Vs[static_cast<size_t>(j) * hd + d]);
                o[0][u] = fmaf(pj.x, v, o[0][u]);
                o[1][u] = fmaf(pj.y, v, o[1][u]);
                o[2][u] = fmaf(pj.z, v, o[2][u]);
                o[3][u] = fmaf(pj.w, v, o[3][u]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int row = row0 + r;
        if (row >= N) break;
        __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * N + row) * H * hd + static_cast<size_t>(h) * hd;
#pragma unroll
        for (int u = 0; u < UMAX; ++u) {
            const int d = lane + 32 * u;
            if (d < hd) dst[d] = __float2bfloat16_rn(o[r][u]);
        }
    }
}
}  // namespace ag

int attention_generic(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_out, int batch, int n_tokens, int heads,
                      int head_dim, float scale, cudaStream_t stream) {
    using namespace ag;
    VTC_REQUIRE(qkv && out, VTC_ERR_ARG, "attention_generic: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0, VTC_ERR_SHAPE, "attention_generic: bad shape");
    VTC_REQUIRE(head_dim >= 16 && head_dim % 16 == 0 && head_dim <= 32 * UMAX, VTC_ERR_SHAPE, "attention_generic: head_dim %d (multiple of 16 up to %d)",
                head_dim, 32 * UMAX);
    VTC_REQUIRE(n_tokens <= 32 * TMAX, VTC_ERR_SHAPE, "attention_generic: %d tokens > %d", n_tokens, 32 * TMAX);
    VTC_REQUIRE(heads <= 65535 && batch <= 65535, VTC_ERR_SHAPE, "attention_generic: grid limits");
    VTC_REQUIRE(scale > 0.f, VTC_ERR_ARG, "attention_generic: scale must be positive");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    Params p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.key_bias = key_bias;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.cls_rows = cls_rows;
    p.attn = attn_out;
    p.B = batch;
    p.N = n_tokens;
    p.H = heads;
    p.hd = head_dim;
    p.npad = (n_tokens + 31) & ~31;
    p.scale = scale;
    const size_t smem = smem_bytes(p.npad, head_dim);
    VTC_REQUIRE(smem <= 232448, VTC_ERR_SHAPE, "attention_generic: %zu bytes of shared memory needed", smem);
    static size_t configured = 0;
    if (smem > configured) {
        VTC_CUDA(cudaFuncSetAttribute(attention_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = smem;
    }
    dim3 grid(cdiv(n_tokens, QT), heads, batch);
    attention_generic_kernel<<<grid, WARPS * 32, smem, stream>>>(p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

}  // namespace vtc

extern "C" {
int vtc_attention_generic(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int32_t batch, int32_t n_tokens,
                          int32_t heads, int32_t head_dim, float scale, void* stream) {
    return vtc::attention_generic(qkv, key_bias, out, cls_rows, attn, batch, n_tokens, heads, head_dim, scale, static_cast<cudaStream_t>(stream));
}
}
