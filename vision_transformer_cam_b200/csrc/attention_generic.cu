// Attention for head dimensions other than 64 (vit_model.py:113-137 with head_dim = embed_dim / num_heads: ViT-H/14 has
// 1280 / 16 = 80).  The tcgen05 kernels (attention_cs / attention_kv / attention.cu) are built around one 128-byte swizzle
// atom per Q / K / V row, i.e. exactly 64 bf16; this kernel is the general-shape path of the same operator, with every
// output of vtc_attention: O, the CLS query row and, on request, the full P.  ViT-H is not a benchmark configuration of the
// reference (whose forward cannot run it at all: 197 tokens and 12 heads are hard-coded); the point is the
// complete factory surface at a reasonable speed, so it uses the warp-level tensor-core path (mma.sync m16n8k16, bf16
// operands, fp32 accumulation) that works for any multiple of 16, not a second tcgen05 pipeline.
//
// One CTA = one (image, head): K and V of the whole sequence are staged once ([Npad][hd+8] bf16 each; the +8 makes every
// fragment load bank-conflict free; two CTAs fit an SM at ViT-H's shape) with the key bias [Npad]; its 8 warps take the
// 16-row query groups round robin, Q fragments straight from global memory.  V stays row-major: the B fragments of P V come
// out of ldmatrix.trans.  Per query group two sweeps over 64-key blocks, both recomputing S = Q K^T on the tensor cores:
//   sweep 1: running row maximum and row sum (online, fp32);
//   sweep 2: P = exp(S - m) / sum -> optional outputs (full P, CLS row) -> bf16 A fragments straight from the accumulator
//            registers -> O += P V.
#include "common.cuh"
#include "ops.h"

namespace vtc {

namespace ag {
constexpr int WARPS = 8;
constexpr int KBLK = 64;          // keys per block
constexpr int MAXKS = 8;          // head_dim / 16 <= 8

struct Params {
    const __nv_bfloat16* qkv;   // [B,N,3,H,hd]
    const float* key_bias;      // [B,N] or null
    __nv_bfloat16* out;         // [B,N,H*hd]
    float* cls_rows;            // [B,H,N] or null
    float* attn;                // [B,H,N,N] or null
    int B, N, H, hd;
    int npad;                   // N rounded up to KBLK
    float scale;
};

__host__ __device__ inline size_t smem_bytes(int npad, int hd) { return static_cast<size_t>(2) * npad * (hd + 8) * 2 + static_cast<size_t>(npad) * 4; }

// four 8x8 b16 matrices, transposed on the way to the registers (B fragments of P V from row-major V)
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// S block (16 rows x 64 keys) of this warp: 8 key tiles x 4 accumulators; scaled, masked, padding keys at -inf.
template <int KS>
__device__ __forceinline__ void score_block(float (&s)[8][4], const uint32_t (&qa)[MAXKS][4], const __nv_bfloat16* Ks, int kp, int key0, int N,
                                            float scale, const float* kbs, bool fg0, bool fg1, int g, int t) {
#pragma unroll
    for (int jt = 0; jt < 8; ++jt) {
        s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
        const __nv_bfloat16* krow = Ks + static_cast<size_t>(key0 + jt * 8 + g) * kp + 2 * t;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + ks * 16);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8);
            mma_bf16_16816(s[jt], qa[ks], b0, b1);
        }
        const int j = key0 + jt * 8 + 2 * t;
        const float kb0 = kbs ? kbs[j] : 0.f, kb1 = kbs ? kbs[j + 1] : 0.f;
        s[jt][0] = j < N ? fmaf(s[jt][0], scale, fg0 ? kb0 : 0.f) : -INFINITY;
        s[jt][1] = j + 1 < N ? fmaf(s[jt][1], scale, fg0 ? kb1 : 0.f) : -INFINITY;
        s[jt][2] = j < N ? fmaf(s[jt][2], scale, fg1 ? kb0 : 0.f) : -INFINITY;
        s[jt][3] = j + 1 < N ? fmaf(s[jt][3], scale, fg1 ? kb1 : 0.f) : -INFINITY;
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

template <int KS>      // head_dim = 16 * KS
__global__ void __launch_bounds__(WARPS * 32, 2) attention_generic_kernel(const Params p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int HD = 16 * KS;
    constexpr int NT = HD / 8;                                   // output tiles of 8 features (even)
    constexpr int kp = HD + 8;                                   // K / V row pitch (bf16)
    const int N = p.N, H = p.H, npad = p.npad;
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* Vs = Ks + static_cast<size_t>(npad) * kp;
    float* kbs = reinterpret_cast<float*>(Vs + static_cast<size_t>(npad) * kp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.y, h = blockIdx.x;
    const size_t tok_stride = static_cast<size_t>(3) * H * HD;
    const __nv_bfloat16* base = p.qkv + static_cast<size_t>(b) * N * tok_stride + static_cast<size_t>(h) * HD;

    // ---- stage K, V (all keys of this head) and the key bias
    constexpr int V8 = HD / 8;                                   // 16-byte vectors per row
    for (int i = threadIdx.x; i < npad * V8; i += blockDim.x) {
        const int n = i / V8, c = i - n * V8;
        uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
        if (n < N) {
            const __nv_bfloat16* tok = base + static_cast<size_t>(n) * tok_stride;
            kv = *reinterpret_cast<const uint4*>(tok + static_cast<size_t>(H) * HD + c * 8);
            vv = *reinterpret_cast<const uint4*>(tok + static_cast<size_t>(2) * H * HD + c * 8);
        }
        *reinterpret_cast<uint4*>(Ks + static_cast<size_t>(n) * kp + c * 8) = kv;
        *reinterpret_cast<uint4*>(Vs + static_cast<size_t>(n) * kp + c * 8) = vv;
    }
    const bool has_bias = p.key_bias != nullptr;
    if (has_bias)
        for (int i = threadIdx.x; i < npad; i += blockDim.x) kbs[i] = i < N ? p.key_bias[static_cast<size_t>(b) * N + i] : 0.f;
    __syncthreads();
    const float* kbp = has_bias ? kbs : nullptr;
    const int nblk = npad / KBLK;
    // ldmatrix source of this lane: matrices 0..3 = (keys +0 / +8) x (features +0 / +8)
    const int lm_key = (lane & 7) + ((lane & 8) ? 8 : 0), lm_feat = (lane & 16) ? 8 : 0;

    for (int rg = warp; rg * 16 < N; rg += WARPS) {
        const int row0 = rg * 16 + g, row1 = row0 + 8;            // this thread's two query rows
        const bool ok0 = row0 < N, ok1 = row1 < N;
        // foreground rows take the key bias; a background row's uniform -100 is softmax-invariant (vit_model.py:348-361)
        const bool fg0 = has_bias && (!ok0 || kbs[row0] == 0.f), fg1 = has_bias && (!ok1 || kbs[row1] == 0.f);
        // Q fragments of the group's 16 rows (rows past the sequence: zeros)
        uint32_t qa[MAXKS][4];
        {
            const __nv_bfloat16* qr0 = base + static_cast<size_t>(row0) * tok_stride + 2 * t;
            const __nv_bfloat16* qr1 = base + static_cast<size_t>(row1) * tok_stride + 2 * t;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                qa[ks][0] = ok0 ? *reinterpret_cast<const uint32_t*>(qr0 + ks * 16) : 0u;
                qa[ks][1] = ok1 ? *reinterpret_cast<const uint32_t*>(qr1 + ks * 16) : 0u;
                qa[ks][2] = ok0 ? *reinterpret_cast<const uint32_t*>(qr0 + ks * 16 + 8) : 0u;
                qa[ks][3] = ok1 ? *reinterpret_cast<const uint32_t*>(qr1 + ks * 16 + 8) : 0u;
            }
        }
        float s[8][4];

        // ---- sweep 1: row maximum and row sum
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        for (int kb = 0; kb < nblk; ++kb) {
            score_block<KS>(s, qa, Ks, kp, kb * KBLK, N, p.scale, kbp, fg0, fg1, g, t);
            float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
            for (int jt = 0; jt < 8; ++jt) {
                bm0 = fmaxf(bm0, fmaxf(s[jt][0], s[jt][1]));
                bm1 = fmaxf(bm1, fmaxf(s[jt][2], s[jt][3]));
            }
            const float n0 = fmaxf(m0, quad_max(bm0)), n1 = fmaxf(m1, quad_max(bm1));      // key 0 is never masked: finite from block 0 on
            l0 *= __expf(m0 - n0);
            l1 *= __expf(m1 - n1);
            m0 = n0;
            m1 = n1;
#pragma unroll
            for (int jt = 0; jt < 8; ++jt) {
                l0 += __expf(s[jt][0] - m0) + __expf(s[jt][1] - m0);
                l1 += __expf(s[jt][2] - m1) + __expf(s[jt][3] - m1);
            }
        }
        const float inv0 = 1.0f / quad_sum(l0), inv1 = 1.0f / quad_sum(l1);

        // ---- sweep 2: P, outputs, O += P V
        float o[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
        float* attn0 = (p.attn && ok0) ? p.attn + ((static_cast<size_t>(b) * H + h) * N + row0) * N : nullptr;
        float* attn1 = (p.attn && ok1) ? p.attn + ((static_cast<size_t>(b) * H + h) * N + row1) * N : nullptr;
        float* cls = (p.cls_rows && row0 == 0) ? p.cls_rows + (static_cast<size_t>(b) * H + h) * N : nullptr;
        for (int kb = 0; kb < nblk; ++kb) {
            score_block<KS>(s, qa, Ks, kp, kb * KBLK, N, p.scale, kbp, fg0, fg1, g, t);
#pragma unroll
            for (int jt = 0; jt < 8; ++jt) {
                s[jt][0] = __expf(s[jt][0] - m0) * inv0;
                s[jt][1] = __expf(s[jt][1] - m0) * inv0;
                s[jt][2] = __expf(s[jt][2] - m1) * inv1;
                s[jt][3] = __expf(s[jt][3] - m1) * inv1;
                const int j = kb * KBLK + jt * 8 + 2 * t;
                if (attn0) {
                    if (j < N) attn0[j] = s[jt][0];
                    if (j + 1 < N) attn0[j + 1] = s[jt][1];
                }
                if (attn1) {
                    if (j < N) attn1[j] = s[jt][2];
                    if (j + 1 < N) attn1[j + 1] = s[jt][3];
                }
                if (cls) {
                    if (j < N) cls[j] = s[jt][0];
                    if (j + 1 < N) cls[j + 1] = s[jt][1];
                }
            }
#pragma unroll
            for (int kk = 0; kk < KBLK / 16; ++kk) {             // 16 keys per k-step = two score tiles
                uint32_t pa[4];
                pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
                pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
                pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
                const __nv_bfloat16* vrow = Vs + static_cast<size_t>(kb * KBLK + kk * 16 + lm_key) * kp + lm_feat;
#pragma unroll
                for (int nt = 0; nt < NT; nt += 2) {
                    uint32_t vb[4];                               // (b0, b1) of feature tile nt, (b0, b1) of tile nt + 1
                    ldsm_x4_trans(vb, vrow + nt * 8);
                    mma_bf16_16816(o[nt], pa, vb[0], vb[1]);
                    mma_bf16_16816(o[nt + 1], pa, vb[2], vb[3]);
                }
            }
        }
        if (ok0) {
            __nv_bfloat16* out0 = p.out + (static_cast<size_t>(b) * N + row0) * H * HD + static_cast<size_t>(h) * HD + 2 * t;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<uint32_t*>(out0 + nt * 8) = pack_bf16x2(o[nt][0], o[nt][1]);
        }
        if (ok1) {
            __nv_bfloat16* out1 = p.out + (static_cast<size_t>(b) * N + row1) * H * HD + static_cast<size_t>(h) * HD + 2 * t;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<uint32_t*>(out1 + nt * 8) = pack_bf16x2(o[nt][2], o[nt][3]);
        }
    }
}

template <int KS>
static int launch(const Params& p, size_t smem, cudaStream_t stream) {
    static SmemOptIn optin;
    int rc_ = optin.ensure(reinterpret_cast<const void*>(attention_generic_kernel<KS>), smem);
    if (rc_ != VTC_OK) return rc_;
    dim3 grid(p.H, p.B);
    attention_generic_kernel<KS><<<grid, WARPS * 32, smem, stream>>>(p);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}
}  // namespace ag

int attention_generic(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_out, int batch, int n_tokens, int heads,
                      int head_dim, float scale, cudaStream_t stream) {
    using namespace ag;
    VTC_REQUIRE(qkv && out, VTC_ERR_ARG, "attention_generic: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 0, VTC_ERR_SHAPE, "attention_generic: bad shape");
    VTC_REQUIRE(head_dim >= 16 && head_dim % 16 == 0 && head_dim <= 16 * MAXKS, VTC_ERR_SHAPE, "attention_generic: head_dim %d (multiple of 16 up to %d)",
                head_dim, 16 * MAXKS);
    VTC_REQUIRE(n_tokens <= kAttentionGenericMaxTokens, VTC_ERR_SHAPE, "attention_generic: %d tokens > %d", n_tokens, kAttentionGenericMaxTokens);
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
    p.npad = cdiv(n_tokens, KBLK) * KBLK;
    p.scale = scale;
    const size_t smem = smem_bytes(p.npad, head_dim);
    VTC_REQUIRE(smem <= 232448, VTC_ERR_SHAPE, "attention_generic: %zu bytes of shared memory needed", smem);
    switch (head_dim / 16) {
        case 1: return launch<1>(p, smem, stream);
        case 2: return launch<2>(p, smem, stream);
        case 3: return launch<3>(p, smem, stream);
        case 4: return launch<4>(p, smem, stream);
        case 5: return launch<5>(p, smem, stream);
        case 6: return launch<6>(p, smem, stream);
        case 7: return launch<7>(p, smem, stream);
        default: return launch<8>(p, smem, stream);
    }
}

}  // namespace vtc

extern "C" {
int vtc_attention_generic(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int32_t batch, int32_t n_tokens,
                          int32_t heads, int32_t head_dim, float scale, void* stream) {
    return vtc::attention_generic(qkv, key_bias, out, cls_rows, attn, batch, n_tokens, heads, head_dim, scale, static_cast<cudaStream_t>(stream));
}
}
