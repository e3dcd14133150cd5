// Small latency-bound kernels around the CLS attention row:
//   cls_stat    head-mean CLS row -> renormalised patch map + batch-global max   (vit_model.py:329-335, 366-372)
//   cls_mask    background decision + additive key bias for the next block     (vit_model.py:335-361)
//   topk_heads  top-16 patches, token gather, head1, final LayerNorm on the CLS row, head  (vit_model.py:374-422)
// None of them materialises a [B,N,N] tensor; everything is fp32.
#include "common.cuh"
#include "ops.h"

namespace vtc {

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < (blockDim.x >> 5)) ? red[lane] : 0.f;
    return warp_sum(t);
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < (blockDim.x >> 5)) ? red[lane] : -INFINITY;
    return warp_max(t);
}

// grid = B, block = 256
__global__ void cls_stat_kernel(const float* __restrict__ cls_rows, float* __restrict__ cls_map, float* __restrict__ gmax, int H, int N) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const float* src = cls_rows + static_cast<size_t>(b) * H * N;
    const float invH = 1.0f / H;
    float part = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += src[h * N + j];
        part += s * invH;
    }
    const float rowsum = block_sum(part, red) + 1.0f;    // + identity on the CLS diagonal (vit_model.py:331-333)
    float mx = 0.f;
    for (int j = threadIdx.x + 1; j < N; j += blockDim.x) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += src[h * N + j];
        const float v = (s * invH) / rowsum;
        cls_map[static_cast<size_t>(b) * (N - 1) + (j - 1)] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    if (threadIdx.x == 0 && gmax != nullptr) atomicMax(reinterpret_cast<int*>(gmax), __float_as_int(mx));   // values are >= 0
}

int cls_stat(const float* cls_rows, float* cls_map, float* gmax, int batch, int heads, int n_tokens, cudaStream_t stream) {
    VTC_REQUIRE(cls_rows && cls_map, VTC_ERR_ARG, "cls_stat: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 1, VTC_ERR_SHAPE, "cls_stat: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    cls_stat_kernel<<<batch, 256, 0, stream>>>(cls_rows, cls_map, gmax, heads, n_tokens);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// grid = B, block = 256
__global__ void cls_mask_kernel(const float* __restrict__ cls_map, const float* __restrict__ gmax, const uint8_t* __restrict__ forced,
                                float thresh, int per_image, uint8_t* __restrict__ bg, float* __restrict__ key_bias, int N) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const int P = N - 1;
    const float* m = cls_map + static_cast<size_t>(b) * P;
    float mx;
    if (per_image) {
        float v = 0.f;
        for (int j = threadIdx.x; j < P; j += blockDim.x) v = fmaxf(v, m[j]);
        mx = block_max(v, red);
    } else {
        mx = *gmax;
    }
    if (threadIdx.x == 0 && key_bias != nullptr) key_bias[static_cast<size_t>(b) * N] = 0.f;
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
        uint8_t isbg;
        if (forced != nullptr) isbg = forced[static_cast<size_t>(b) * P + j] != 0;
        else isbg = (m[j] / mx) < thresh;                         // torch.lt(mask_14 / max, 0.25)
        if (bg != nullptr) bg[static_cast<size_t>(b) * P + j] = isbg;
        if (key_bias != nullptr) key_bias[static_cast<size_t>(b) * N + 1 + j] = isbg ? -100.0f : 0.0f;
    }
}

int cls_mask(const float* cls_map, const float* gmax, const uint8_t* forced_bg, float thresh, int per_image, uint8_t* bg,
             float* key_bias, int batch, int n_tokens, cudaStream_t stream) {
    VTC_REQUIRE(cls_map && (per_image || gmax), VTC_ERR_ARG, "cls_mask: null pointer");
    VTC_REQUIRE(batch > 0 && n_tokens > 1, VTC_ERR_SHAPE, "cls_mask: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    cls_mask_kernel<<<batch, 256, 0, stream>>>(cls_map, gmax, forced_bg, thresh, per_image, bg, key_bias, n_tokens);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// cls_stat + cls_mask in ONE launch (the forward, layers >= mask_from: 8 launches less per ViT-B forward).  Every block runs
// the cls_stat body for its image, publishes its maximum (atomicMax) and draws a ticket; the batch-global maximum of
// vit_model.py:335 is complete when all B tickets are drawn, so every block waits for that (all blocks of the grid are
// co-resident: the host checks B against the occupancy limit and otherwise launches the two kernels) and then thresholds
// its own image.  `ticket`: one zeroed unsigned int per launch (the forward zeroes one per layer next to gmax).
__global__ void __launch_bounds__(256) cls_stat_mask_kernel(const float* __restrict__ cls_rows, float* __restrict__ cls_map, float* __restrict__ gmax,
                                                            const uint8_t* __restrict__ forced, float thresh, int per_image, uint8_t* __restrict__ bg,
                                                            float* __restrict__ key_bias, unsigned int* __restrict__ ticket, int B, int H, int N,
                                                            uint8_t* __restrict__ aug, AugLayout al, float inv_scale) {
    __shared__ float red[32];
    __shared__ float mapv[2048];       // this image's map (n_tokens <= 2049, checked by the host)
    const int b = blockIdx.x;
    const int P = N - 1;
    const float* src = cls_rows + static_cast<size_t>(b) * H * N;
    const float invH = 1.0f / H;
    float part = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += src[h * N + j];
        if (j > 0) mapv[j - 1] = s * invH;
        part += s * invH;
    }
    const float rowsum = block_sum(part, red) + 1.0f;    // + identity on the CLS diagonal (vit_model.py:331-333)
    float mx = 0.f;
    for (int j = threadIdx.x; j < P; j += blockDim.x) {   // (block_sum synchronised the block: every mapv entry is visible)
        const float v = mapv[j] / rowsum;
        mapv[j] = v;
        cls_map[static_cast<size_t>(b) * P + j] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    if (!per_image) {
        if (threadIdx.x == 0) {
            atomicMax(reinterpret_cast<int*>(gmax), __float_as_int(mx));      // values are >= 0
            __threadfence();
            atomicAdd(ticket, 1u);
            while (*reinterpret_cast<volatile unsigned int*>(ticket) < static_cast<unsigned int>(B)) __nanosleep(64);
            __threadfence();
            red[0] = *reinterpret_cast<volatile float*>(gmax);
        }
        __syncthreads();
        mx = red[0];
    } else if (threadIdx.x == 0) {
        atomicMax(reinterpret_cast<int*>(gmax), __float_as_int(mx));          // topk_heads may still want the batch-global value
    }
    // mask operands of the fast attention kernel (ops.h: AugLayout): K_aug[key] = bias / scale, Q_aug[row] = [row not masked];
    // only the first bf16 of a row's first 16-byte half ever changes (the buffer was zeroed at the start of the forward)
    uint8_t* ak = aug ? aug + static_cast<size_t>(b) * al.per_image : nullptr;
    uint8_t* aq = aug ? ak + al.nb * al.k_block_bytes : nullptr;
    auto put_aug = [&](int token, bool masked) {
        const int jb = token / al.KB, rk = token - jb * al.KB;
        *reinterpret_cast<uint32_t*>(ak + jb * al.k_block_bytes + (rk >> 3) * 256 + (rk & 7) * 16) = pack_bf16x2(masked ? -100.0f * inv_scale : 0.f, 0.f);
        const int qt = token >> 7, rq = token & 127;
        *reinterpret_cast<uint32_t*>(aq + qt * al.q_tile_bytes + (rq >> 3) * 256 + (rq & 7) * 16) = pack_bf16x2(masked ? 0.f : 1.0f, 0.f);
    };
    if (threadIdx.x == 0) {
        key_bias[static_cast<size_t>(b) * N] = 0.f;
        if (aug) put_aug(0, false);
    }
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
        const size_t e = static_cast<size_t>(b) * P + j;
        const uint8_t isbg = forced ? (forced[e] != 0) : ((mapv[j] / mx) < thresh);          // torch.lt(mask_14 / max, 0.25)
        if (bg != nullptr) bg[e] = isbg;
        key_bias[static_cast<size_t>(b) * N + 1 + j] = isbg ? -100.0f : 0.0f;
        if (aug) put_aug(1 + j, isbg != 0);
    }
}

// largest batch the one-launch kernel serves on the current device: every block of the grid must be resident at the same time
int cls_stat_mask_capacity() {
    static std::atomic<int> per_sm[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int bps = (dev >= 0 && dev < 64) ? per_sm[dev].load(std::memory_order_relaxed) : 0;
    if (bps == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, cls_stat_mask_kernel, 256, 0) != cudaSuccess) return 0;
        if (dev >= 0 && dev < 64) per_sm[dev].store(bps, std::memory_order_relaxed);
    }
    return bps * device_sm_count();
}

int cls_stat_mask(const float* cls_rows, float* cls_map, float* gmax, const uint8_t* forced_bg, float thresh, int per_image, uint8_t* bg,
                  float* key_bias, unsigned int* ticket, int batch, int heads, int n_tokens, cudaStream_t stream, void* aug, float inv_scale) {
    VTC_REQUIRE(cls_rows && cls_map && gmax && key_bias && ticket, VTC_ERR_ARG, "cls_stat_mask: null pointer");
    VTC_REQUIRE(batch > 0 && heads > 0 && n_tokens > 1, VTC_ERR_SHAPE, "cls_stat_mask: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int capacity = cls_stat_mask_capacity();
    VTC_REQUIRE(!aug || (n_tokens - 1 <= 2048 && batch <= capacity), VTC_ERR_SHAPE, "cls_stat_mask: mask operands need the one-launch path (batch %d)", batch);
    if (n_tokens - 1 > 2048 || batch > capacity) {
        if ((rc = cls_stat(cls_rows, cls_map, gmax, batch, heads, n_tokens, stream)) != VTC_OK) return rc;
        return cls_mask(cls_map, gmax, forced_bg, thresh, per_image, bg, key_bias, batch, n_tokens, stream);
    }
    cls_stat_mask_kernel<<<batch, 256, 0, stream>>>(cls_rows, cls_map, gmax, forced_bg, thresh, per_image, bg, key_bias, ticket, batch, heads, n_tokens,
                                                    static_cast<uint8_t*>(aug), attention_aug_layout(n_tokens), inv_scale);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// grid = B, block = 256.  Dynamic smem: [P] map copy + [D] mean token + [D] normalised CLS row + [R] pre-logits.
// The map is divided by its maximum first (batch-global `*gmax`, or the image's own when gmax == nullptr) exactly as
// vit_model.py:372 does before torch.topk (:377): the fp32 quotients, not the raw values, decide order and ties.
__global__ void topk_heads_kernel(const HeadParams hp, const float* __restrict__ tokens, const float* __restrict__ cls_map,
                                  const float* __restrict__ gmax, const int32_t* __restrict__ forced_topk, float* __restrict__ logits,
                                  float* __restrict__ hwp_logits, float* __restrict__ hwp_tokens, int32_t* __restrict__ topk_idx) {
    extern __shared__ float sm[];
    __shared__ float red[32];
    __shared__ int idx_s[64];
    const int b = blockIdx.x;
    const int N = hp.n_tokens, P = N - 1, D = hp.dim, K = hp.topk, C = hp.classes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* mapv = sm;            // [P]
    float* meanv = mapv + P;     // [D]
    float* xn = meanv + D;       // [D]
    float* feat = xn + D;        // [R] (only with pre_logits)

    if (forced_topk != nullptr) {
        if (threadIdx.x < K) idx_s[threadIdx.x] = min(max(forced_topk[b * K + threadIdx.x], 0), P - 1);
    } else {
        float mx;
        if (gmax != nullptr) {
            mx = *gmax;
        } else {
            float v = 0.f;
            for (int j = threadIdx.x; j < P; j += blockDim.x) v = fmaxf(v, cls_map[static_cast<size_t>(b) * P + j]);
            mx = block_max(v, red);
        }
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
            const float v = cls_map[static_cast<size_t>(b) * P + j] / mx;
            mapv[j] = (v != v) ? INFINITY : v;            // torch.topk ranks NaN above every number
        }
        __syncthreads();
        if (warp == 0) {
            // K rounds of warp arg-max, descending, ties to the smaller index (torch.topk, vit_model.py:377); taken entries are
            // marked with a NaN so that -inf values remain selectable and an index is never returned twice
            for (int k = 0; k < K; ++k) {
                float best = 0.f;
                int bi = 0x7fffffff;
                for (int j = lane; j < P; j += 32) {
                    const float v = mapv[j];
                    if (v == v && (bi == 0x7fffffff || v > best)) { best = v; bi = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
                }
                if (lane == 0) { idx_s[k] = bi; mapv[bi] = __int_as_float(0x7fc00000); }      // K <= P: a candidate always exists
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (topk_idx != nullptr && threadIdx.x < K) topk_idx[b * K + threadIdx.x] = idx_s[threadIdx.x];

    // gather the K tokens (block-L output, pre final norm) and their mean; the CLS row goes to shared memory once
    const float* tok = tokens + static_cast<size_t>(b) * N * D;
    const float invK = 1.0f / K;
    for (int d = threadIdx.x; d < D; d += blockDim.x) xn[d] = tok[d];
    for (int d4 = threadIdx.x; d4 < (D >> 2); d4 += blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 v = ldg_f4(tok + static_cast<size_t>(1 + idx_s[k]) * D + 4 * d4);
            st_f4(hwp_tokens + (static_cast<size_t>(b) * K + k) * D + 4 * d4, v);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        meanv[4 * d4] = s.x * invK; meanv[4 * d4 + 1] = s.y * invK; meanv[4 * d4 + 2] = s.z * invK; meanv[4 * d4 + 3] = s.w * invK;
    }
    for (int d = (D & ~3) + threadIdx.x; d < D; d += blockDim.x) {      // D not a multiple of 4 (never for a ViT)
        float s = 0.f;
        for (int k = 0; k < K; ++k) {
            const float v = tok[static_cast<size_t>(1 + idx_s[k]) * D + d];
            hwp_tokens[(static_cast<size_t>(b) * K + k) * D + d] = v;
            s += v;
        }
        meanv[d] = s * invK;
    }
    __syncthreads();
    // final LayerNorm of the CLS row only (vit_model.py:402 normalises all rows; only row 0 is consumed :406)
    float part = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) part += xn[d];
    const float mean = block_sum(part, red) / D;
    part = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { const float c = xn[d] - mean; part += c * c; }
    const float rstd = rsqrtf(block_sum(part, red) / D + hp.eps);
    for (int d = threadIdx.x; d < D; d += blockDim.x) xn[d] = (xn[d] - mean) * rstd * hp.norm_w[d] + hp.norm_b[d];
    __syncthreads();

    const float* fin = xn;
    int F = D;
    if (hp.pre_w != nullptr) {   // pre_logits = tanh(fc(x)) (vit_model.py:267-273)
        for (int r = warp; r < hp.rep; r += nwarps) {
            float s = 0.f;
            for (int d = lane; d < D; d += 32) s += hp.pre_w[static_cast<size_t>(r) * D + d] * xn[d];
            s = warp_sum(s);
            if (lane == 0) feat[r] = tanhf(s + hp.pre_b[r]);
        }
        __syncthreads();
        fin = feat;
        F = hp.rep;
    }
    const bool vec4 = ((F | D | P) & 3) == 0;          // 16-byte loads: the weight rows and the shared-memory vectors are 16-byte aligned then
    for (int c = warp; c < C; c += nwarps) {
        float s = 0.f, s1 = 0.f;
        if (vec4) {
            const float4* w4 = reinterpret_cast<const float4*>(hp.head_w + static_cast<size_t>(c) * F);
            const float4* v4 = reinterpret_cast<const float4*>(hp.head1_w + static_cast<size_t>(c) * D);
            for (int d = lane; d < (F >> 2); d += 32) {
                const float4 a = __ldg(w4 + d), x = reinterpret_cast<const float4*>(fin)[d];
                s += (a.x * x.x + a.y * x.y) + (a.z * x.z + a.w * x.w);
            }
            for (int d = lane; d < (D >> 2); d += 32) {
                const float4 a = __ldg(v4 + d), x = reinterpret_cast<const float4*>(meanv)[d];
                s1 += (a.x * x.x + a.y * x.y) + (a.z * x.z + a.w * x.w);
            }
        } else {
            for (int d = lane; d < F; d += 32) s += hp.head_w[static_cast<size_t>(c) * F + d] * fin[d];
            for (int d = lane; d < D; d += 32) s1 += hp.head1_w[static_cast<size_t>(c) * D + d] * meanv[d];
        }
        s = warp_sum(s);
        s1 = warp_sum(s1);
        if (lane == 0) {
            logits[b * C + c] = s + hp.head_b[c];
            hwp_logits[b * C + c] = s1 + hp.head1_b[c];
        }
    }
}

int topk_heads(const HeadParams& hp, const float* tokens, const float* cls_map, const float* gmax, const int32_t* forced_topk, float* logits,
               float* hwp_logits, float* hwp_tokens, int32_t* topk_idx, int batch, cudaStream_t stream) {
    VTC_REQUIRE(tokens && cls_map && logits && hwp_logits && hwp_tokens, VTC_ERR_ARG, "topk_heads: null pointer");
    VTC_REQUIRE(hp.norm_w && hp.norm_b && hp.head_w && hp.head_b && hp.head1_w && hp.head1_b, VTC_ERR_ARG, "topk_heads: missing weights");
    VTC_REQUIRE(hp.topk > 0 && hp.topk <= 64 && hp.topk <= hp.n_tokens - 1, VTC_ERR_SHAPE, "topk_heads: topk=%d", hp.topk);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const size_t smem = sizeof(float) * (static_cast<size_t>(hp.n_tokens - 1) + 2 * hp.dim + (hp.pre_w ? hp.rep : 0));
    VTC_REQUIRE(smem <= 48 * 1024, VTC_ERR_SHAPE, "topk_heads: %zu bytes of smem", smem);
    topk_heads_kernel<<<batch, 256, smem, stream>>>(hp, tokens, cls_map, gmax, forced_topk, logits, hwp_logits, hwp_tokens, topk_idx);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

}  // namespace vtc

extern "C" {
int vtc_cls_stat(const float* cls_rows, float* cls_map, float* gmax, int32_t batch, int32_t heads, int32_t n_tokens, void* stream) {
    return vtc::cls_stat(cls_rows, cls_map, gmax, batch, heads, n_tokens, static_cast<cudaStream_t>(stream));
}
size_t vtc_attention_mask_operand_bytes(int32_t n_tokens) { return n_tokens > 0 ? vtc::attention_aug_layout(n_tokens).per_image : 0; }
int vtc_cls_stat_mask(const float* cls_rows, float* cls_map, float* gmax, const uint8_t* forced_bg, float thresh, int32_t per_image, uint8_t* bg,
                      float* key_bias, uint32_t* ticket, void* mask_operands, float inv_scale, int32_t batch, int32_t heads, int32_t n_tokens, void* stream) {
    return vtc::cls_stat_mask(cls_rows, cls_map, gmax, forced_bg, thresh, per_image, bg, key_bias, ticket, batch, heads, n_tokens,
                              static_cast<cudaStream_t>(stream), mask_operands, inv_scale);
}
int vtc_cls_mask(const float* cls_map, const float* gmax, const uint8_t* forced_bg, float thresh, int32_t per_image, uint8_t* bg,
                 float* key_bias, int32_t batch, int32_t n_tokens, void* stream) {
    return vtc::cls_mask(cls_map, gmax, forced_bg, thresh, per_image, bg, key_bias, batch, n_tokens, static_cast<cudaStream_t>(stream));
}
}
