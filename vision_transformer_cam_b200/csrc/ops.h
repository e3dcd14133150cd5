// Internal C++ entry points of the kernels (the extern "C" wrappers in each .cu and the model forward call these).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vtc {

// split != 0 ("fp32 mode"): A [M,2K] and W [N,2K] hold (hi | lo) bf16 halves, bf16 outputs are written as [M,2N] halves
int gemm_bf16(const void* A, const void* W, const float* bias, const float* residual, const float* pos, void* out, int M,
              int N, int K, int epilogue, int tokens, cudaStream_t stream, int split = 0, int reverse = 0);
// LayerNorm-fused GEMMs of the bf16 forward (gemm.cu): stats = [M][D/128][2] partial (sum, sum of squares) per 128-column slice
int gemm_resid_ln(const void* A, const void* W, const float* bias, const float* residual, float* out, void* out_bf16, float* stats, int M, int N,
                  int K, cudaStream_t stream, int reverse = 0);
int gemm_lnfold(const void* A, const void* W, const float* c, const float* g, const float* stats, float eps, void* out, int M, int N, int K,
                int gelu, cudaStream_t stream, int reverse = 0);
int residual_prep(const float* x, void* xb, float* stats, int rows, int dim, cudaStream_t stream);
int fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* g, float* c, int N, int K, cudaStream_t stream);
int cast_bf16(const float* src, void* dst, size_t n, cudaStream_t stream);
// fp32 [rows,cols] -> (hi | lo) bf16 halves [rows, 2*cols], x ~= hi + lo (16 mantissa bits)
int split_bf16(const float* src, void* dst, size_t rows, size_t cols, cudaStream_t stream);
int patchify(const float* x, void* patches, int batch, int in_c, int img, int patch, cudaStream_t stream, int split = 0);
// decoded u8 HWC images [B,S,S,3] -> normalised bf16 patch matrix; mean / std are HOST arrays of 3 floats
// any patch size, K padded with zeros to kp columns (bf16 mode, fp32 input)
int patchify_generic(const float* x, void* patches, int batch, int in_c, int img, int patch, int kp, cudaStream_t stream);
int patchify_generic_u8(const uint8_t* x, const float* mean, const float* std, void* patches, int batch, int img, int patch, int kp,
                        cudaStream_t stream);
int cast_bf16_pad(const float* src, void* dst, size_t rows, int cols, int ld, cudaStream_t stream);
int patchify_u8(const uint8_t* x, const float* mean, const float* std, void* patches, int batch, int img, int patch, cudaStream_t stream,
                int split = 0);
int cls_token_rows(const float* cls_token, const float* pos_embed, float* tokens, int batch, int n_tokens, int dim, cudaStream_t stream);
int layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int rows, int dim, float eps, cudaStream_t stream,
                   int split = 0, int reverse = 0);
// Precomputed mask operands of the fast attention kernel (attention_cs.cu applies the reference's additive mask through one
// extra tensor-core K-step with Q_aug[i] = [v_i == 0], K_aug[j] = key_bias[j] / scale).  cls_stat_mask writes them per image in
// the exact shared-memory image of that K-step (no-swizzle core matrices: row r at (r >> 3) * 256 + (r & 7) * 16, second
// 16-byte half 128 bytes further, always zero), so that the attention producer fetches them with two bulk copies per item
// instead of rebuilding them with generic stores for every (head, query tile): image b at aug + b * per_image =
// [nb key blocks x KB rows x 32 B] then [qtiles x 128 rows x 32 B].  The buffer must be zeroed once (padding rows / halves).
struct AugLayout {
    int KB, nb, qtiles;
    size_t k_block_bytes, q_tile_bytes, per_image;
};
constexpr int kAttentionSingleBlockKeysFwd = 208, kAttentionLongBlockKeysFwd = 192;      // == kAttentionSingleBlockKeys / kAttentionLongBlockKeys below
inline AugLayout attention_aug_layout(int n_tokens) {
    AugLayout a{};
    if (n_tokens <= kAttentionSingleBlockKeysFwd) { a.nb = 1; a.KB = (n_tokens + 15) & ~15; }
    else { a.nb = (n_tokens + kAttentionLongBlockKeysFwd - 1) / kAttentionLongBlockKeysFwd; a.KB = (((n_tokens + a.nb - 1) / a.nb) + 31) & ~31; }
    a.qtiles = (n_tokens + 127) / 128;
    a.k_block_bytes = static_cast<size_t>(a.KB) * 32;
    a.q_tile_bytes = 128 * 32;
    a.per_image = a.nb * a.k_block_bytes + a.qtiles * a.q_tile_bytes;
    return a;
}
int attention(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int batch, int n_tokens, int heads,
              float scale, cudaStream_t stream, int reverse = 0, const void* aug = nullptr);
// Attention + head mean of P (the rollout's input, predict.py:189-190) without a [B,H,N,N] fp32 round trip, any n_tokens the
// fast kernel serves: it stores the bf16 exponentials it feeds to P V ("packed P", see attention_cs.cu) into `scratch`
// (attention_mean_scratch_bytes), head_mean_packed reduces them over the heads into attn_mean [B,N,N].
struct PackedP {
    void* e;        // bf16 [B,H,N,ld]
    float* mtab;    // [B,H,N,ld/32]: reference maximum of every 32-key chunk (log2 domain); null for one key block
    float* mfin;    // [B,H,N]: final row maximum; null for one key block
    float* einv;    // [B,H,N]: 1 / row sum (relative to mfin)
};
constexpr int kAttentionFusedMeanMaxTokens = 2048;
constexpr int kAttentionSingleBlockKeys = 208;     // attention_cs: whole sequence in one key block up to here,
constexpr int kAttentionLongBlockKeys = 192;       // blocks of at most this many keys beyond
// keys per stored row of the packed P = the key blocking of attention_cs rounded to whole 32-key chunks
inline int attention_packed_ld(int n_tokens) {
    if (n_tokens <= kAttentionSingleBlockKeys) return ((n_tokens + 31) / 32) * 32;
    const int nb = (n_tokens + kAttentionLongBlockKeys - 1) / kAttentionLongBlockKeys;
    return nb * ((((n_tokens + nb - 1) / nb) + 31) / 32 * 32);
}
size_t attention_mean_scratch_bytes(int batch, int n_tokens, int heads);
int attention_mean(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* scratch, size_t scratch_bytes,
                   int batch, int n_tokens, int heads, float scale, cudaStream_t stream, int reverse = 0);
int head_mean_packed(const PackedP& packed, float* mean, int batch, int heads, int n_tokens, int ld, cudaStream_t stream);
// "Rollout operand": the head mean of one layer as the rollout kernel streams it (bf16, half the bytes of the
// fp32 [B,N,N] matrix): bf16 [B,N,ldr], ldr = rollout_operand_ld(N); a row holds N values, zero padding, and in its last
// four bytes the fp32 sum of the N ROUNDED values (so that (Pbar + I) / rowsum is normalised exactly as stored).
inline int rollout_operand_ld(int n_tokens) { return (n_tokens + 2 + 7) / 8 * 8; }
int head_mean_packed_operand(const PackedP& packed, void* operand, int batch, int heads, int n_tokens, int ld, cudaStream_t stream);
// attention_mean writing the rollout operand instead of (mean == nullptr) or next to the fp32 mean
int attention_mean_operand(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* operand, void* scratch,
                           size_t scratch_bytes, int batch, int n_tokens, int heads, float scale, cudaStream_t stream, int reverse = 0,
                           const void* aug = nullptr);
// fp32 [B,N,N] head mean -> rollout operand (the path of the full-P / fp32-mode / general-shape forwards)
int rollout_operand_from_mean(const float* mean, void* operand, int batch, int n_tokens, cudaStream_t stream);
// r <- e0^T A_{L-1} ... A_0 from `layers` rollout operands [layers,B,N,ldr]; row [B,N-1]
int rollout_operand(const void* operands, float* row, int layers, int batch, int n_tokens, cudaStream_t stream);
int rollout(const float* attn_mean, float* row, int layers, int batch, int n_tokens, cudaStream_t stream);
// KV-blocked kernel (attention_kv.cu): any n_tokens <= 2048; split = (hi, lo) bf16 operands: qkv [B,N,2,3,H,64], out [B*N,2,H*64]
int attention_kv(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int batch, int n_tokens, int heads,
                 float scale, bool split, cudaStream_t stream, int reverse = 0);
// column-split pipelined kernel (attention_cs.cu): the fast path when the full P is not requested, any n_tokens <= 2048
// packed: optional packed-P output (row stride attention_packed_ld(N)) for attention_mean
int attention_cs(const void* qkv, const float* key_bias, void* out, float* cls_rows, int batch, int n_tokens, int heads, float scale,
                 cudaStream_t stream, int reverse = 0, const PackedP* packed = nullptr, const void* aug = nullptr);
// general-shape path (attention_generic.cu): head_dim a multiple of 16 up to 128, n_tokens <= 320; same outputs as `attention`
int attention_generic(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int batch, int n_tokens, int heads,
                      int head_dim, float scale, cudaStream_t stream);
constexpr int kAttentionGenericMaxTokens = 320;
unsigned long long* attention_trace_buffer();   // debug: device buffer for %globaltimer stamps (vtc_debug_set_attention_trace) or null
int head_mean(const float* attn, float* mean, int batch, int heads, int n_tokens, cudaStream_t stream);
int cls_stat(const float* cls_rows, float* cls_map, float* gmax, int batch, int heads, int n_tokens, cudaStream_t stream);
int cls_mask(const float* cls_map, const float* gmax, const uint8_t* forced_bg, float thresh, int per_image, uint8_t* bg,
             float* key_bias, int batch, int n_tokens, cudaStream_t stream);
// both in one launch (the forward); `ticket`: a zeroed unsigned int that the kernel leaves zeroed
int cls_stat_mask_capacity();      // largest batch of the one-launch kernel on the current device
// aug (optional): the attention mask operands of every image (attention_aug_layout), inv_scale = 1 / attention scale
int cls_stat_mask(const float* cls_rows, float* cls_map, float* gmax, const uint8_t* forced_bg, float thresh, int per_image, uint8_t* bg,
                  float* key_bias, unsigned int* ticket, int batch, int heads, int n_tokens, cudaStream_t stream, void* aug = nullptr,
                  float inv_scale = 0.f);

struct HeadParams {
    const float* norm_w; const float* norm_b;
    const float* pre_w;  const float* pre_b;
    const float* head_w; const float* head_b;
    const float* head1_w; const float* head1_b;
    int dim, rep, classes, topk, n_tokens;
    float eps;
};
int topk_heads(const HeadParams& hp, const float* tokens, const float* cls_map, const float* gmax, const int32_t* forced_topk, float* logits,
               float* hwp_logits, float* hwp_tokens, int32_t* topk_idx, int batch, cudaStream_t stream);

}  // namespace vtc
