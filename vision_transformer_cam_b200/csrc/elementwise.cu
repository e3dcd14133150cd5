// Bandwidth-bound kernels of the forward: fp32->bf16 cast, patchify (im2col of the k=s=p conv), CLS token rows,
// LayerNorm.  All use 16-byte vector accesses and are coalesced on the side that moves the most bytes.
#include "common.cuh"
#include "ops.h"

namespace vtc {

// ---- cast ------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    const size_t n8 = n / 8;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = ld_stream_f4(src + i * 8);
        const float4 b = ld_stream_f4(src + i * 8 + 4);
        st_u4(dst + i * 8, make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w)));
    }
    if (blockIdx.x == 0 && threadIdx.x < n - n8 * 8) dst[n8 * 8 + threadIdx.x] = __float2bfloat16_rn(src[n8 * 8 + threadIdx.x]);
}

int cast_bf16(const float* src, void* dst, size_t n, cudaStream_t stream) {
    VTC_REQUIRE(src && dst, VTC_ERR_ARG, "cast: null pointer");
    if (n == 0) return VTC_OK;
    VTC_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, VTC_ERR_ARG,
                "cast: pointers must be 16-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const size_t n8 = n / 8;
    size_t blocks = (n8 + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    cast_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// fp32 [rows, cols] -> bf16 [rows, ld] with the columns cols .. ld-1 zeroed (weights of a GEMM whose K is padded, e.g. the
// 3*14*14 = 588 -> 640 patch-embedding contraction of ViT-H/14)
__global__ void cast_bf16_pad_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t rows, int cols, int ld) {
    const size_t total = rows * ld;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t r = i / ld;
        const int c = static_cast<int>(i - r * ld);
        dst[i] = __float2bfloat16_rn(c < cols ? src[r * cols + c] : 0.f);
    }
}

int cast_bf16_pad(const float* src, void* dst, size_t rows, int cols, int ld, cudaStream_t stream) {
    VTC_REQUIRE(src && dst, VTC_ERR_ARG, "cast_pad: null pointer");
    VTC_REQUIRE(rows > 0 && cols > 0 && ld >= cols, VTC_ERR_SHAPE, "cast_pad: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    size_t blocks = (rows * ld + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    cast_bf16_pad_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), rows, cols, ld);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- split (fp32 mode) ------------------------------------------------------------------------------
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits.  Rows keep their halves side by side, [rows, 2*cols].
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
    hi = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    lo = make_uint4(pack_bf16x2(a.x - __uint_as_float(hi.x << 16), a.y - __uint_as_float(hi.x & 0xffff0000u)),
                    pack_bf16x2(a.z - __uint_as_float(hi.y << 16), a.w - __uint_as_float(hi.y & 0xffff0000u)),
                    pack_bf16x2(b.x - __uint_as_float(hi.z << 16), b.y - __uint_as_float(hi.z & 0xffff0000u)),
                    pack_bf16x2(b.z - __uint_as_float(hi.w << 16), b.w - __uint_as_float(hi.w & 0xffff0000u)));
}

__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t rows, size_t cols) {
    const size_t c8 = cols / 8;
    const size_t total = rows * c8;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t r = i / c8, c = (i - r * c8) * 8;
        const float4 a = ld_stream_f4(src + r * cols + c);
        const float4 b = ld_stream_f4(src + r * cols + c + 4);
        uint4 hi, lo;
        split8(a, b, hi, lo);
        st_u4(dst + r * 2 * cols + c, hi);
        st_u4(dst + r * 2 * cols + cols + c, lo);
    }
}

int split_bf16(const float* src, void* dst, size_t rows, size_t cols, cudaStream_t stream) {
    VTC_REQUIRE(src && dst, VTC_ERR_ARG, "split: null pointer");
    VTC_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0, VTC_ERR_SHAPE, "split: cols must be a positive multiple of 8");
    VTC_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, VTC_ERR_ARG,
                "split: pointers must be 16-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const size_t total = rows * (cols / 8);
    size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    split_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), rows, cols);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- patchify --------------------------------------------------------------------------------------
// One thread moves 8 consecutive pixels of one image row: reads are fully coalesced along x (the fp32 side is
// 2/3 of the bytes); each 16-byte bf16 store lands in patch row (b,py,px) at k = c*p*p + kh*p + kw.
template <bool SPLIT>
__global__ void patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int in_c, int S, int p, size_t total8) {
    const int g = S / p;
    const int kdim = in_c * p * p;
    const int xg_per_row = S / 8;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8; i += stride) {
        const int xg = static_cast<int>(i % xg_per_row);
        size_t r = i / xg_per_row;
        const int y = static_cast<int>(r % S);
        r /= S;
        const int c = static_cast<int>(r % in_c);
        const size_t b = r / in_c;
        const float4 a0 = ld_stream_f4(x + i * 8);
        const float4 a1 = ld_stream_f4(x + i * 8 + 4);
        const int x0 = xg * 8;
        const int py = y / p, kh = y - py * p;
        const int px = x0 / p, kw = x0 - px * p;
        const size_t row = (b * g + py) * g + px;
        __nv_bfloat16* dst = out + row * kdim * (SPLIT ? 2 : 1) + (c * p + kh) * p + kw;
        uint4 hi, lo;
        split8(a0, a1, hi, lo);
        st_u4(dst, hi);
        if (SPLIT) st_u4(dst + kdim, lo);
    }
}

int patchify(const float* x, void* patches, int batch, int in_c, int img, int patch, cudaStream_t stream, int split) {
    VTC_REQUIRE(x && patches, VTC_ERR_ARG, "patchify: null pointer");
    VTC_REQUIRE(batch > 0 && in_c > 0 && img > 0 && patch > 0, VTC_ERR_SHAPE, "patchify: bad shape");
    VTC_REQUIRE(img % patch == 0 && patch % 8 == 0, VTC_ERR_SHAPE, "patchify: img %d / patch %d unsupported (patch %% 8 == 0 required)", img, patch);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const size_t total8 = static_cast<size_t>(batch) * in_c * img * img / 8;
    size_t blocks = (total8 + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    if (split) patchify_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(patches), in_c, img, patch, total8);
    else patchify_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(patches), in_c, img, patch, total8);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// Any patch size (ViT-H/14: 14 is not a multiple of the 8-pixel vectors above), K padded to `kp` columns with zeros: one
// thread per element of the patch matrix.  Reads are 4-byte gathers along image rows (p consecutive pixels per run).
__global__ void patchify_generic_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int in_c, int S, int p, int kp, size_t total) {
    const int g = S / p;
    const int kdim = in_c * p * p;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t row = i / kp;
        const int k = static_cast<int>(i - row * kp);
        float v = 0.f;
        if (k < kdim) {
            const int c = k / (p * p), rem = k - c * p * p, kh = rem / p, kw = rem - kh * p;
            const size_t b = row / (static_cast<size_t>(g) * g);
            const int pr = static_cast<int>(row - b * g * g), py = pr / g, px = pr - py * g;
            v = x[((b * in_c + c) * S + (py * p + kh)) * S + px * p + kw];
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

int patchify_generic(const float* x, void* patches, int batch, int in_c, int img, int patch, int kp, cudaStream_t stream) {
    VTC_REQUIRE(x && patches, VTC_ERR_ARG, "patchify: null pointer");
    VTC_REQUIRE(batch > 0 && in_c > 0 && img > 0 && patch > 0 && img % patch == 0 && kp >= in_c * patch * patch, VTC_ERR_SHAPE, "patchify: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const size_t total = static_cast<size_t>(batch) * (img / patch) * (img / patch) * kp;
    size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    patchify_generic_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(patches), in_c, img, patch, kp, total);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- patchify from decoded uint8 images -----------------------------------------------
// The reference feeds PIL images through ToTensor (u8 / 255) and Normalize ((x - mean) / std) on the host
// (predict.py:72-75, validate.py:80-84) and ships fp32 NCHW to the GPU.  Here the decoded u8 HWC pixels are shipped (4x
// fewer bytes over PCIe and out of HBM) and normalised in the im2col pass, with the same two IEEE fp32 operations, so the
// bf16 patch matrix is bit-identical to patchify(Normalize(ToTensor(img))).
struct NormParams {
    float mean[3];
    float std[3];
};

template <bool SPLIT>
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ out, int S, int p, size_t total8, NormParams nm) {
    // 256 values per channel: the two IEEE operations are evaluated once per table entry (bit-identical to doing them per
    // pixel; the per-pixel divisions made this kernel issue-bound and slower than its fp32 counterpart: 53 vs 37 us)
    __shared__ float lut[3][256];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8;
        lut[c][i & 255] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i & 255), 255.0f), nm.mean[c]), nm.std[c]);
    }
    __syncthreads();
    const int g = S / p;
    const int kdim = 3 * p * p;
    const int xg_per_row = S / 8;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8; i += stride) {
        // i indexes 8 consecutive pixels (24 bytes, 8-byte aligned) of image row y
        const int xg = static_cast<int>(i % xg_per_row);
        size_t r = i / xg_per_row;
        const int y = static_cast<int>(r % S);
        const size_t b = r / S;
        const uint2* src = reinterpret_cast<const uint2*>(x + i * 24);
        const uint2 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
        const uint32_t w[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
        const int x0 = xg * 8;
        const int py = y / p, kh = y - py * p;
        const int px = x0 / p, kw = x0 - px * p;
        const size_t row = (b * g + py) * g + px;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int byte = q * 3 + c;
                const uint32_t u = (w[byte >> 2] >> ((byte & 3) * 8)) & 0xffu;
                v[q] = lut[c][u];
            }
            uint4 hi, lo;
            split8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
            __nv_bfloat16* dst = out + row * kdim * (SPLIT ? 2 : 1) + (c * p + kh) * p + kw;
            st_u4(dst, hi);
            if (SPLIT) st_u4(dst + kdim, lo);
        }
    }
}

// uint8 HWC input, any patch size, K padded to kp (the general-shape counterpart of patchify_u8: same two fp32 operations)
__global__ void patchify_generic_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ out, int S, int p, int kp, size_t total,
                                           NormParams nm) {
    const int g = S / p;
    const int kdim = 3 * p * p;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t row = i / kp;
        const int k = static_cast<int>(i - row * kp);
        float v = 0.f;
        if (k < kdim) {
            const int c = k / (p * p), rem = k - c * p * p, kh = rem / p, kw = rem - kh * p;
            const size_t b = row / (static_cast<size_t>(g) * g);
            const int pr = static_cast<int>(row - b * g * g), py = pr / g, px = pr - py * g;
            const float u = static_cast<float>(x[((b * S + (py * p + kh)) * S + px * p + kw) * 3 + c]);
            v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), nm.mean[c]), nm.std[c]);
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

int patchify_generic_u8(const uint8_t* x, const float* mean, const float* std, void* patches, int batch, int img, int patch, int kp,
                        cudaStream_t stream) {
    VTC_REQUIRE(x && mean && std && patches, VTC_ERR_ARG, "patchify_u8: null pointer");
    VTC_REQUIRE(batch > 0 && img > 0 && patch > 0 && img % patch == 0 && kp >= 3 * patch * patch, VTC_ERR_SHAPE, "patchify_u8: bad shape");
    for (int c = 0; c < 3; ++c) VTC_REQUIRE(std[c] != 0.f, VTC_ERR_ARG, "patchify_u8: std[%d] is zero", c);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    NormParams nm{{mean[0], mean[1], mean[2]}, {std[0], std[1], std[2]}};
    const size_t total = static_cast<size_t>(batch) * (img / patch) * (img / patch) * kp;
    size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    patchify_generic_u8_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(patches), img, patch, kp, total, nm);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

int patchify_u8(const uint8_t* x, const float* mean, const float* std, void* patches, int batch, int img, int patch, cudaStream_t stream, int split) {
    VTC_REQUIRE(x && mean && std && patches, VTC_ERR_ARG, "patchify_u8: null pointer");
    VTC_REQUIRE(batch > 0 && img > 0 && patch > 0, VTC_ERR_SHAPE, "patchify_u8: bad shape");
    VTC_REQUIRE(img % patch == 0 && patch % 8 == 0, VTC_ERR_SHAPE, "patchify_u8: img %d / patch %d unsupported (patch %% 8 == 0 required)", img, patch);
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0, VTC_ERR_ARG, "patchify_u8: image pointer must be 8-byte aligned");
    for (int c = 0; c < 3; ++c) VTC_REQUIRE(std[c] != 0.f, VTC_ERR_ARG, "patchify_u8: std[%d] is zero", c);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    NormParams nm{{mean[0], mean[1], mean[2]}, {std[0], std[1], std[2]}};
    const size_t total8 = static_cast<size_t>(batch) * img * img / 8;
    size_t blocks = (total8 + 255) / 256;
    const size_t cap = static_cast<size_t>(device_sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    if (split) patchify_u8_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(patches), img, patch, total8, nm);
    else patchify_u8_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(patches), img, patch, total8, nm);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- CLS token rows ---------------------------------------------------------------------------------
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ tokens,
                                int n_tokens, int dim) {
    float* dst = tokens + static_cast<size_t>(blockIdx.x) * n_tokens * dim;
    for (int i = threadIdx.x; i < dim / 4; i += blockDim.x) {
        const float4 a = ldg_f4(cls + 4 * i);
        const float4 b = ldg_f4(pos + 4 * i);
        st_f4(dst + 4 * i, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
    }
}

int cls_token_rows(const float* cls_token, const float* pos_embed, float* tokens, int batch, int n_tokens, int dim, cudaStream_t stream) {
    VTC_REQUIRE(cls_token && pos_embed && tokens, VTC_ERR_ARG, "cls_token_rows: null pointer");
    VTC_REQUIRE(batch > 0 && dim % 4 == 0, VTC_ERR_SHAPE, "cls_token_rows: bad shape");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    cls_rows_kernel<<<batch, 256, 0, stream>>>(cls_token, pos_embed, tokens, n_tokens, dim);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- LayerNorm ---------------------------------------------------------------------------------------
// One warp per row, the row lives in registers (NV float4 per lane), two-pass mean / variance in fp32 exactly
// like ATen's (sum of squared deviations), bf16 output.  8 rows per 256-thread CTA.
template <int NV, bool SPLIT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows,
                                                        float eps, int reverse) {
    constexpr int D = NV * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int row = blockIdx.x * 8 + warp;
    if (row >= rows) return;
    if (reverse) row = rows - 1 - row;       // last rows first: they are the ones the previous kernel left in L2
    const float* src = x + static_cast<size_t>(row) * D;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(src + (lane + 32 * i) * 4);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    __nv_bfloat16* dst = y + static_cast<size_t>(row) * D * (SPLIT ? 2 : 1);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        const float4 g = ldg_f4(gamma + c);
        const float4 b = ldg_f4(beta + c);
        const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
        const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
        const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
        const uint2 hi = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        *reinterpret_cast<uint2*>(dst + c) = hi;
        if (SPLIT)
            *reinterpret_cast<uint2*>(dst + D + c) =
                make_uint2(pack_bf16x2(o0 - __uint_as_float(hi.x << 16), o1 - __uint_as_float(hi.x & 0xffff0000u)),
                           pack_bf16x2(o2 - __uint_as_float(hi.y << 16), o3 - __uint_as_float(hi.y & 0xffff0000u)));
    }
}

template <int NV>
static void launch_ln(int grid, cudaStream_t stream, const float* x, const float* gamma, const float* beta, __nv_bfloat16* out, int rows, float eps,
                      int split, int reverse) {
    if (split) layernorm_kernel<NV, true><<<grid, 256, 0, stream>>>(x, gamma, beta, out, rows, eps, reverse);
    else layernorm_kernel<NV, false><<<grid, 256, 0, stream>>>(x, gamma, beta, out, rows, eps, reverse);
}

int layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int rows, int dim, float eps, cudaStream_t stream,
                   int split, int reverse) {
    VTC_REQUIRE(x && gamma && beta && y, VTC_ERR_ARG, "layernorm: null pointer");
    VTC_REQUIRE(rows > 0, VTC_ERR_SHAPE, "layernorm: rows=%d", rows);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int grid = cdiv(rows, 8);
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(y);
    switch (dim) {
        case 256: launch_ln<2>(grid, stream, x, gamma, beta, out, rows, eps, split, reverse); break;
        case 384: launch_ln<3>(grid, stream, x, gamma, beta, out, rows, eps, split, reverse); break;
        case 512: launch_ln<4>(grid, stream, x, gamma, beta, out, rows, eps, split, reverse); break;
        case 768: launch_ln<6>(grid, stream, x, gamma, beta, out, rows, eps, split, reverse); break;
        case 1024: launch_ln<8>(grid, stream, x, gamma, beta, out, rows, eps, split, reverse); break;
        case 1280: launch_ln<10>(grid, stream, x, gamma, beta, out, rows, eps, split, reverse); break;
        default:
            set_last_error("layernorm: dim %d unsupported (256/384/512/768/1024/1280)", dim);
            return VTC_ERR_SHAPE;
    }
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// ---- LayerNorm fusion helpers (bf16 mode, see gemm.cu) -------------------------------------------------------------------
// residual_prep: what the EPI_RESID_LN epilogue leaves behind, for a residual stream that did not come out of that epilogue
// (the patch-embedding output feeding block 0): bf16 copy + per-row partial (sum, sum of squares) of every 128-column slice.
template <int NV>
__global__ void __launch_bounds__(256) residual_prep_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float* __restrict__ stats,
                                                            int rows) {
    constexpr int D = NV * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float* src = x + static_cast<size_t>(row) * D;
    __nv_bfloat16* dst = xb + static_cast<size_t>(row) * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {          // slice i = columns [128 i, 128 i + 128): one float4 per lane
        const float4 v = *reinterpret_cast<const float4*>(src + i * 128 + lane * 4);
        *reinterpret_cast<uint2*>(dst + i * 128 + lane * 4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        const float s1 = warp_sum((v.x + v.y) + (v.z + v.w));
        const float s2 = warp_sum(fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w))));
        if (lane == 0) *reinterpret_cast<float2*>(stats + (static_cast<size_t>(row) * NV + i) * 2) = make_float2(s1, s2);
    }
}

int residual_prep(const float* x, void* xb, float* stats, int rows, int dim, cudaStream_t stream) {
    VTC_REQUIRE(x && xb && stats, VTC_ERR_ARG, "residual_prep: null pointer");
    VTC_REQUIRE(rows > 0, VTC_ERR_SHAPE, "residual_prep: rows=%d", rows);
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    const int grid = cdiv(rows, 8);
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(xb);
    switch (dim) {
        case 256: residual_prep_kernel<2><<<grid, 256, 0, stream>>>(x, out, stats, rows); break;
        case 512: residual_prep_kernel<4><<<grid, 256, 0, stream>>>(x, out, stats, rows); break;
        case 768: residual_prep_kernel<6><<<grid, 256, 0, stream>>>(x, out, stats, rows); break;
        case 1024: residual_prep_kernel<8><<<grid, 256, 0, stream>>>(x, out, stats, rows); break;
        case 1280: residual_prep_kernel<10><<<grid, 256, 0, stream>>>(x, out, stats, rows); break;
        default:
            set_last_error("residual_prep: dim %d unsupported (256/512/768/1024/1280)", dim);
            return VTC_ERR_SHAPE;
    }
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

// fold_ln: W'[n,:] = bf16(gamma * W[n,:]), g[n] = sum_k W'[n,k] (of the ROUNDED values: rstd (t.W'^T - mean g) is then exactly
// rstd ((t - mean).W'^T)), c[n] = bias[n] + sum_k beta[k] W[n,k].  One warp per output row.
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ Wf, float* __restrict__ g,
                                                      float* __restrict__ c, int N, int K) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + warp;
    if (n >= N) return;
    const float* src = W + static_cast<size_t>(n) * K;
    __nv_bfloat16* dst = Wf + static_cast<size_t>(n) * K;
    float sg = 0.f, sc = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
        const float4 w = ldg_f4(src + k), ga = ldg_f4(gamma + k), be = ldg_f4(beta + k);
        const uint32_t p0 = pack_bf16x2(w.x * ga.x, w.y * ga.y), p1 = pack_bf16x2(w.z * ga.z, w.w * ga.w);
        *reinterpret_cast<uint2*>(dst + k) = make_uint2(p0, p1);
        sg += (__uint_as_float(p0 << 16) + __uint_as_float(p0 & 0xffff0000u)) + (__uint_as_float(p1 << 16) + __uint_as_float(p1 & 0xffff0000u));
        sc = fmaf(w.x, be.x, fmaf(w.y, be.y, fmaf(w.z, be.z, fmaf(w.w, be.w, sc))));
    }
    sg = warp_sum(sg);
    sc = warp_sum(sc);
    if (lane == 0) {
        g[n] = sg;
        c[n] = bias[n] + sc;
    }
}

int fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* g, float* c, int N, int K, cudaStream_t stream) {
    VTC_REQUIRE(W && gamma && beta && bias && Wf && g && c, VTC_ERR_ARG, "fold_ln: null pointer");
    VTC_REQUIRE(N > 0 && K > 0 && K % 128 == 0, VTC_ERR_SHAPE, "fold_ln: K must be a multiple of 128");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    fold_ln_kernel<<<cdiv(N, 8), 256, 0, stream>>>(W, gamma, beta, bias, static_cast<__nv_bfloat16*>(Wf), g, c, N, K);
    VTC_CHECK_LAUNCH();
    return VTC_OK;
}

}  // namespace vtc

extern "C" {
int vtc_residual_prep(const float* x, void* xb, float* stats, int32_t rows, int32_t dim, void* stream) {
    return vtc::residual_prep(x, xb, stats, rows, dim, static_cast<cudaStream_t>(stream));
}
int vtc_fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* g, float* c, int32_t N, int32_t K,
                void* stream) {
    return vtc::fold_ln(W, gamma, beta, bias, Wf, g, c, N, K, static_cast<cudaStream_t>(stream));
}
int vtc_cast_bf16(const float* src, void* dst, size_t n, void* stream) {
    return vtc::cast_bf16(src, dst, n, static_cast<cudaStream_t>(stream));
}
int vtc_split_bf16(const float* src, void* dst, size_t rows, size_t cols, void* stream) {
    return vtc::split_bf16(src, dst, rows, cols, static_cast<cudaStream_t>(stream));
}
int vtc_patchify(const float* x, void* patches, int32_t batch, int32_t in_c, int32_t img, int32_t patch, void* stream) {
    return vtc::patchify(x, patches, batch, in_c, img, patch, static_cast<cudaStream_t>(stream), 0);
}
int vtc_patchify_split(const float* x, void* patches, int32_t batch, int32_t in_c, int32_t img, int32_t patch, void* stream) {
    return vtc::patchify(x, patches, batch, in_c, img, patch, static_cast<cudaStream_t>(stream), 1);
}
int vtc_patchify_u8(const uint8_t* x, const float* mean, const float* std, void* patches, int32_t batch, int32_t img, int32_t patch,
                     int32_t split, void* stream) {
    return vtc::patchify_u8(x, mean, std, patches, batch, img, patch, static_cast<cudaStream_t>(stream), split);
}
int vtc_layernorm_split(const float* x, const float* gamma, const float* beta, void* y, int32_t rows, int32_t dim, float eps,
                        void* stream) {
    return vtc::layernorm_bf16(x, gamma, beta, y, rows, dim, eps, static_cast<cudaStream_t>(stream), 1, 0);
}
int vtc_cls_token_rows(const float* cls_token, const float* pos_embed, float* tokens, int32_t batch, int32_t n_tokens, int32_t dim,
                       void* stream) {
    return vtc::cls_token_rows(cls_token, pos_embed, tokens, batch, n_tokens, dim, static_cast<cudaStream_t>(stream));
}
int vtc_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int32_t rows, int32_t dim, float eps,
                       void* stream) {
    return vtc::layernorm_bf16(x, gamma, beta, y, rows, dim, eps, static_cast<cudaStream_t>(stream), 0, 0);
}
}
