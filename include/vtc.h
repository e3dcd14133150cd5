/*
 * vtc.h -- C-ABI of libvtc.so: the B200 (sm_100a) ViT forward + CAM / attention-rollout hot path.
 *
 * The reference (Jingfeng-Tang/vision_transformer_cam) has no FFI: its "operator API" for this path is the
 * Python surface of vit_model.py plus the inline post-processing of predict.py / validate.py.  This header
 * is what a ctypes binding of that path binds (INTEGRATION.md shows the stub).  Each entry point cites the
 * reference lines it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - every function returns int: VTC_OK (0) or a negative VTC_ERR_*; `vtc_last_error()` (thread local)
 *     describes the last failure.  Nothing throws, nothing prints.
 *   - all tensor pointers are DEVICE pointers owned by the caller (torch allocations), dense row-major,
 *     borrowed until the enqueued work has run.  The library never allocates or frees device memory and
 *     never synchronises: work is enqueued on the `stream` argument (a cudaStream_t passed as void*).
 *   - sm_100a only.  There is no CPU path and no other-architecture path: on anything else the first
 *     enqueue returns VTC_ERR_ARCH.
 *   - shapes: B images, N = 1 + (img/patch)^2 tokens, D embed_dim, H heads, hd = D/H, L depth, C classes,
 *     P = N-1 patches, g = img/patch, K = topk (16).
 */
#ifndef VTC_H_
#define VTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTC_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define VTC_API __attribute__((visibility("default")))
#else
#define VTC_API
#endif

enum {
    VTC_OK = 0,
    VTC_ERR_ARG = -1,       /* null pointer / bad enum / inconsistent option */
    VTC_ERR_SHAPE = -2,     /* unsupported shape (e.g. embed_dim not a multiple of 128) */
    VTC_ERR_ARCH = -3,      /* device is not compute capability 10.x */
    VTC_ERR_CUDA = -4,      /* a CUDA runtime / driver call failed */
    VTC_ERR_WORKSPACE = -5  /* workspace or packed-weight buffer too small / misaligned */
};

/* ------------------------------------------------------------------------------------------------
 * Model description: VisionTransformer.__init__ arguments (vit_model.py:215-219) that shape the forward.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vtc_config {
    int32_t img_size;            /* 224 | 384 | 448 ... (square, vit_model.py:53-56) */
    int32_t patch_size;          /* 16 */
    int32_t in_c;                /* 3 */
    int32_t num_classes;         /* C */
    int32_t embed_dim;           /* D, multiple of 128, head_dim must be 64 */
    int32_t depth;               /* L */
    int32_t num_heads;           /* H */
    int32_t mlp_hidden;          /* int(D * mlp_ratio), multiple of 256 */
    int32_t representation_size; /* 0 = pre_logits is Identity (vit_model.py:267-276) */
    int32_t mask_from;           /* 4: first layer whose CLS attention builds a mask (vit_model.py:118,325) */
    float   mask_thresh;         /* 0.25 (vit_model.py:339) */
    int32_t topk;                /* 16 (vit_model.py:377) */
    float   ln_eps;              /* 1e-6 (vit_model.py:244) */
} vtc_config;

/* fp32 parameters, state_dict names in comments.  Device pointers. */
typedef struct vtc_layer_weights {
    const float* norm1_w; const float* norm1_b;   /* blocks.i.norm1.{weight,bias}      [D] */
    const float* qkv_w;   const float* qkv_b;     /* blocks.i.attn.qkv.{weight,bias}   [3D,D],[3D] */
    const float* proj_w;  const float* proj_b;    /* blocks.i.attn.proj.{weight,bias}  [D,D],[D] */
    const float* norm2_w; const float* norm2_b;   /* blocks.i.norm2.{weight,bias}      [D] */
    const float* fc1_w;   const float* fc1_b;     /* blocks.i.mlp.fc1.{weight,bias}    [4D,D],[4D] */
    const float* fc2_w;   const float* fc2_b;     /* blocks.i.mlp.fc2.{weight,bias}    [D,4D],[D] */
} vtc_layer_weights;

typedef struct vtc_weights {
    const float* cls_token;                        /* cls_token  [1,1,D] */
    const float* pos_embed;                        /* pos_embed  [1,N,D] */
    const float* patch_w; const float* patch_b;    /* patch_embed.proj.{weight,bias} [D,in_c,p,p],[D] */
    const float* norm_w;  const float* norm_b;     /* norm.{weight,bias} [D] */
    const float* pre_w;   const float* pre_b;      /* pre_logits.fc.{weight,bias} [R,D],[R] or NULL */
    const float* head_w;  const float* head_b;     /* head.{weight,bias}  [C,R or D],[C] */
    const float* head1_w; const float* head1_b;    /* head1.{weight,bias} [C,R or D],[C] */
    const vtc_layer_weights* layers;               /* host array of `num_layers` entries */
    int32_t num_layers;
} vtc_weights;

/* forward flags */
enum {
    VTC_FWD_MASK_NORM_IMAGE = 1 << 0, /* normalise the CLS map per image instead of the reference's batch-global
                                         max (vit_model.py:335,372) */
    VTC_FWD_FP32_SPLIT      = 1 << 1  /* the "fp32 mode" (logits within 1e-4 of the fp32 reference): must be set iff the
                                         model's precision is VTC_PRECISION_FP32_SPLIT */
};

/* Arithmetic of the dense contractions (vtc_model_set_precision).
 *   VTC_PRECISION_BF16        bf16 operands, fp32 accumulation (default; BASELINE's bf16 mode, logits within 1e-2)
 *   VTC_PRECISION_FP32_SPLIT  every GEMM / attention operand x is carried as a pair of bf16 (hi, lo), x ~= hi + lo
 *                             (16 mantissa bits), and every product is evaluated on the tensor cores as
 *                             hi.hi + lo.hi + hi.lo with fp32 accumulation: 3x the MMA work, ~2^-17 relative error */
enum { VTC_PRECISION_BF16 = 0, VTC_PRECISION_FP32_SPLIT = 1 };

/* Optional teacher forcing of the discrete decisions (tests only; NULL = compute them). */
typedef struct vtc_forcing {
    const uint8_t* bg;       /* [L,B,P] background vectors; used for layers with bg_layer_mask bit set */
    uint32_t bg_layer_mask;  /* bit l set -> layer l's bg vector is taken from `bg` */
    const int32_t* topk_idx; /* [B,K] or NULL */
} vtc_forcing;

/* Forward outputs.  NULL = not requested (except logits / hwp_logits / hwp_tokens which are required). */
typedef struct vtc_outputs {
    float*   logits;      /* [B,C]       head(norm(x)[:,0])                       vit_model.py:402-422 */
    float*   hwp_logits;  /* [B,C]       head1(mean of the K high-weight tokens)  vit_model.py:391-393 */
    float*   hwp_tokens;  /* [B,K,D]     ori_allbs_hw_p_ts                        vit_model.py:381-390 */
    int32_t* topk_idx;    /* [B,K]       patch indices, descending attention      vit_model.py:377 */
    float*   tokens;      /* [Lt,B,N,D]  block outputs X_l (attn_matrix)          vit_model.py:324 */
    int32_t  tokens_layers; /* Lt: how many trailing layers to keep: 1 (last only) .. L */
    float*   cls_rows;    /* [L,B,H,N]   P_l[:,:,0,:]  (CLS query row per head)   vit_model.py:329-334 */
    float*   attn;        /* [La,B,H,N,N] full softmax P_l (attn_weights)         vit_model.py:126-128,323 */
    int32_t  attn_layers; /* La trailing layers kept in `attn` (0 = none) */
    float*   attn_mean;   /* [L,B,N,N]   head-mean P_l (what rollout consumes)    predict.py:189-190 */
    uint8_t* bg;          /* [L,B,P]     background vector built after layer l (0 for l < mask_from) :337-342 */
    float*   cls_map;     /* [L,B,P]     renormalised head-mean CLS row (before /max)  vit_model.py:329-334 */
    float*   rollout;     /* [B,P]       attention rollout row over the last min(L,12) layers, un-normalised: (e0^T A_{L-1} .. A_0)[1:],
                                          A_l = (head-mean P_l + I) / rowsum; the head means stay in the workspace as bf16
                                          "rollout operands" (half the bytes of attn_mean)              predict.py:215-232 */
} vtc_outputs;

typedef struct vtc_model vtc_model;   /* host-side handle: config, packed-weight pointers, TMA descriptors */

/* ---- lifecycle ------------------------------------------------------------------------------- */
VTC_API int vtc_version(void);
VTC_API const char* vtc_last_error(void);
/* VTC_OK iff the current CUDA device is sm_100 class (compute capability 10.x). */
VTC_API int vtc_check_device(void);
/* number of CUDA kernels this library has launched in this process (every vtc_* kernel launch counts once) */
VTC_API uint64_t vtc_launch_count(void);

/* replaces VisionTransformer.__init__ shape bookkeeping (vit_model.py:215-301) */
VTC_API int vtc_model_create(const vtc_config* cfg, vtc_model** out);
VTC_API int vtc_model_destroy(vtc_model* m);
/* Select the arithmetic BEFORE vtc_model_packed_bytes / vtc_model_pack_weights / vtc_workspace_bytes / vtc_forward:
 * it changes the size and layout of the packed weights and of the workspace. */
VTC_API int vtc_model_set_precision(vtc_model* m, int32_t precision);
/* bytes of the packed (bf16, K-major) GEMM weight buffer the caller must provide */
VTC_API size_t vtc_model_packed_bytes(const vtc_model* m);
/* fp32 state_dict tensors -> packed bf16 buffer (enqueued on stream); the fp32 vectors (biases, LayerNorm,
 * cls_token, pos_embed, heads) are referenced in place and must outlive the handle's use.
 * Replaces nothing in the reference (it feeds fp32 nn.Linear weights to ATen); call again after any
 * in-place weight update (load_state_dict). */
VTC_API int vtc_model_pack_weights(vtc_model* m, const vtc_weights* w, void* packed, size_t packed_bytes, void* stream);
/* workspace bytes for a batch of B images (depends on which outputs are requested: pass the same struct) */
VTC_API size_t vtc_workspace_bytes(const vtc_model* m, int32_t batch, const vtc_outputs* outs);

/* ---- the fused path ---------------------------------------------------------------------------
 * VisionTransformer.forward (vit_model.py:303-424): patch embed, L pre-norm blocks with the layer>4
 * background mask, high-weight-patch head, final norm + head.  x: [B,in_c,S,S] fp32 NCHW. */
VTC_API int vtc_forward(vtc_model* m, const float* x, int32_t batch, const vtc_outputs* outs, const vtc_forcing* forcing,
                void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* The same forward fed with decoded images (the step right before the path, predict.py:72-75 /
 * validate.py:80-84): x uint8 [B,S,S,3] HWC on the device, mean / std HOST arrays of 3 floats (Normalize); ToTensor and
 * Normalize are applied inside the patch-matrix kernel, bit-identically to the fp32 path.  4x fewer input bytes. */
VTC_API int vtc_forward_u8(vtc_model* m, const uint8_t* x, const float* mean, const float* std, int32_t batch, const vtc_outputs* outs,
                   const vtc_forcing* forcing, void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* ---- per-kernel timing of the forward (bench.py's roofline numbers) -----------------------------
 * When enabled, vtc_forward brackets every kernel launch with CUDA events on the launch stream;
 * vtc_model_profile_read synchronises on them and ADDS the elapsed milliseconds / launch counts per kind
 * into the caller's arrays (length VTC_PROF_KINDS), then forgets the recorded spans. */
enum {
    VTC_PROF_PATCHIFY = 0, VTC_PROF_GEMM_PATCH = 1, VTC_PROF_LAYERNORM = 2, VTC_PROF_GEMM_QKV = 3, VTC_PROF_ATTENTION = 4,
    VTC_PROF_GEMM_PROJ = 5, VTC_PROF_GEMM_FC1 = 6, VTC_PROF_GEMM_FC2 = 7, VTC_PROF_CLS = 8, VTC_PROF_HEAD_MEAN = 9,
    VTC_PROF_HEADS = 10, VTC_PROF_ROLLOUT = 11, VTC_PROF_KINDS = 12
};
VTC_API int vtc_model_profile(vtc_model* m, int32_t enable);
VTC_API int vtc_model_profile_read(vtc_model* m, float* ms_per_kind, int32_t* launches_per_kind);

/* ---- per-kernel entry points (unit parity; the model forward is built from these) -------------- */
enum { VTC_EPI_BIAS = 0, VTC_EPI_BIAS_GELU = 1, VTC_EPI_BIAS_RESIDUAL = 2, VTC_EPI_PATCH_EMBED = 3 };
/* out[M,Nout] = A[M,K] (bf16) . W[Nout,K]^T (bf16) + bias (+ epilogue).  nn.Linear: vit_model.py:110,138,158,161;
 * conv-as-GEMM: vit_model.py:64,76.
 *   VTC_EPI_BIAS / _GELU : out is bf16 [M,Nout]
 *   VTC_EPI_BIAS_RESIDUAL: out is fp32 [M,Nout] = residual + A.W^T + bias (out may alias residual)
 *   VTC_EPI_PATCH_EMBED  : A rows are (b,p) patches; out is the fp32 token buffer [B,tokens,Nout], row b*tokens+1+p,
 *                          plus pos_embed[1+p] (pos = [tokens,Nout] fp32); vit_model.py:79,308-314
 * Requires K % 64 == 0, Nout % 256 == 0. */
VTC_API int vtc_gemm_bf16(const void* A, const void* W, const float* bias, const float* residual, const float* pos,
                  void* out, int32_t M, int32_t Nout, int32_t K, int32_t epilogue, int32_t tokens, void* stream);

/* The same GEMM on split operands (VTC_PRECISION_FP32_SPLIT): A [M,2K] and W [Nout,2K] hold (hi | lo) bf16 halves per row
 * (vtc_split_bf16); bf16 outputs are written as halves too, [M,2*Nout]; fp32 outputs are unchanged. */
VTC_API int vtc_gemm_split(const void* A, const void* W, const float* bias, const float* residual, const float* pos,
                   void* out, int32_t M, int32_t Nout, int32_t K, int32_t epilogue, int32_t tokens, void* stream);
/* fp32 [rows,cols] -> (hi | lo) bf16 halves [rows,2*cols]: hi = bf16(x), lo = bf16(x - hi); cols % 8 == 0 */
VTC_API int vtc_split_bf16(const float* src, void* dst, size_t rows, size_t cols, void* stream);
/* vtc_patchify / vtc_layernorm_bf16 with split outputs: patches [B*P, 2*in_c*p*p], y [rows, 2*D] */
VTC_API int vtc_patchify_split(const float* x, void* patches, int32_t batch, int32_t in_c, int32_t img, int32_t patch, void* stream);
VTC_API int vtc_layernorm_split(const float* x, const float* gamma, const float* beta, void* y, int32_t rows, int32_t dim,
                        float eps, void* stream);

/* vtc_patchify for decoded images: x uint8 [B,S,S,3] HWC; mean / std: HOST arrays of 3 floats.  Computes
 * ((u8 / 255) - mean) / std with the reference's two fp32 operations (ToTensor + Normalize, predict.py:72-75), so the
 * patch matrix is bit-identical to vtc_patchify of the normalised fp32 tensor.  split != 0: (hi | lo) halves. */
VTC_API int vtc_patchify_u8(const uint8_t* x, const float* mean, const float* std, void* patches, int32_t batch, int32_t img,
                    int32_t patch, int32_t split, void* stream);

/* ---- LayerNorm fused into the GEMMs either side of it (what vtc_forward runs in bf16 mode) ---------------------------
 * The pre-norm block computes x + f(LN(x)) twice (vit_model.py:189-198).  Instead of a LayerNorm kernel per call, the GEMM
 * that updates the residual stream also emits what the next LayerNorm needs, and the GEMM that consumes LN(x) applies the
 * normalisation in its epilogue:
 *   LN(t).W^T + b = rstd (bf16(t).W'^T - mean g) + c,  W' = gamma * W,  g[n] = sum_k W'[n,k],  c[n] = b[n] + sum_k beta[k] W[n,k].
 * Row statistics travel as partial (sum, sum of squares) pairs per 128-column slice: stats [M, D/128, 2] fp32.
 *
 * vtc_gemm_resid_ln: out = residual + A.W^T + bias (fp32, out may alias residual), out_bf16 = bf16(out), stats of out.
 * vtc_gemm_lnfold:   out (bf16) = [GELU](rstd (A.W'^T - mean g) + c) with mean / rstd from `stats` (K = D columns), eps = LN eps.
 * vtc_residual_prep: x fp32 [rows,D] -> bf16 copy + stats (for a residual stream that did not come out of vtc_gemm_resid_ln).
 * vtc_fold_ln:       W fp32 [Nout,K], gamma / beta [K], bias [Nout] -> W' bf16 [Nout,K], g [Nout], c [Nout]. */
VTC_API int vtc_gemm_resid_ln(const void* A, const void* W, const float* bias, const float* residual, float* out, void* out_bf16,
                      float* stats, int32_t M, int32_t Nout, int32_t K, void* stream);
VTC_API int vtc_gemm_lnfold(const void* A, const void* W, const float* c, const float* g, const float* stats, float eps, void* out,
                    int32_t M, int32_t Nout, int32_t K, int32_t gelu, void* stream);
VTC_API int vtc_residual_prep(const float* x, void* xb, float* stats, int32_t rows, int32_t dim, void* stream);
VTC_API int vtc_fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* g, float* c,
                int32_t Nout, int32_t K, void* stream);

/* fp32 -> bf16 (weight packing, also used by tests) */
VTC_API int vtc_cast_bf16(const float* src, void* dst, size_t n, void* stream);
/* NCHW fp32 image -> bf16 patch matrix [B*P, in_c*p*p] (k = c*p*p + kh*p + kw), the im2col of the k=s=p conv
 * (vit_model.py:64,76-79) */
VTC_API int vtc_patchify(const float* x, void* patches, int32_t batch, int32_t in_c, int32_t img, int32_t patch, void* stream);
/* tokens[b,0,:] = cls_token + pos_embed[0] (vit_model.py:308-314) */
VTC_API int vtc_cls_token_rows(const float* cls_token, const float* pos_embed, float* tokens, int32_t batch, int32_t n_tokens,
                       int32_t dim, void* stream);
/* nn.LayerNorm over the last dim (vit_model.py:193,198,402): x fp32 [rows,D] -> y bf16 [rows,D] */
VTC_API int vtc_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int32_t rows, int32_t dim,
                       float eps, void* stream);
/* Attention core (vit_model.py:113-137): qkv bf16 [B,N,3,H,64] -> out bf16 [B,N,H*64];
 * key_bias [B,N] = -100*v (v = background vector incl. the CLS entry 0) or NULL: the reference mask
 * -100*min(v_i+v_j,1) (vit_model.py:118-124,348-361), i.e. key j is biased by key_bias[j] on query rows with v_i = 0
 * and rows with v_i = 1 stay unmasked (their uniform -100 is softmax-invariant);
 * cls_rows [B,H,N] fp32 (P[b,h,0,:]) or NULL; attn [B,H,N,N] fp32 full P or NULL. */
VTC_API int vtc_attention(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int32_t batch,
                  int32_t n_tokens, int32_t heads, float scale, void* stream);
/* vtc_attention + the head mean of P, attn_mean [B,N,N] = mean_h P[b,h] (what the rollout consumes, predict.py:189-190),
 * without materialising [B,H,N,N] fp32: the kernel writes the bf16 exponentials it feeds to P.V and 1/rowsum into
 * `scratch` (256-byte aligned, >= vtc_attention_mean_scratch_bytes), a second kernel reduces them over the heads in a
 * fixed order.  Any n_tokens <= 2048. */
VTC_API size_t vtc_attention_mean_scratch_bytes(int32_t batch, int32_t n_tokens, int32_t heads);
VTC_API int vtc_attention_mean(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* scratch,
                       size_t scratch_bytes, int32_t batch, int32_t n_tokens, int32_t heads, float scale, void* stream);
/* vtc_attention for head dimensions other than 64 (ViT-H/14: 1280 / 16 = 80): qkv [B,N,3,H,head_dim], head_dim a multiple
 * of 16 up to 128, n_tokens <= 320; fp32 arithmetic on bf16 operands, same outputs and mask semantics (vit_model.py:113-137). */
VTC_API int vtc_attention_generic(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int32_t batch,
                          int32_t n_tokens, int32_t heads, int32_t head_dim, float scale, void* stream);
/* The KV-blocked kernel behind vtc_attention for n_tokens > 256 (ViT-B/16-448: 785 tokens, ViT-L/16-384: 577 tokens),
 * callable directly for any n_tokens <= 2048.  Same arguments and outputs as vtc_attention.
 *   split != 0 ("fp32 mode"): operands are (hi, lo) bf16 pairs, x ~= hi + lo: qkv is [B,N,2,3,H,64] (all hi parts of a
 *   token, then all lo parts), out is [B*N,2,H*64]; every product is evaluated as hi.hi + lo.hi + hi.lo in fp32. */
VTC_API int vtc_attention_kv(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn, int32_t batch,
                     int32_t n_tokens, int32_t heads, float scale, int32_t split, void* stream);
/* head mean of P: attn [B,H,N,N] -> mean [B,N,N] (predict.py:189-190) */
VTC_API int vtc_head_mean(const float* attn, float* mean, int32_t batch, int32_t heads, int32_t n_tokens, void* stream);
/* CLS-row statistic (vit_model.py:329-335): cls_rows [B,H,N] -> cls_map [B,P] = ((mean_h + e0)/rowsum)[1:], and
 * gmax[0] = max(gmax[0], max cls_map) (caller zeroes gmax).  */
VTC_API int vtc_cls_stat(const float* cls_rows, float* cls_map, float* gmax, int32_t batch, int32_t heads, int32_t n_tokens,
                 void* stream);
/* background mask (vit_model.py:335-361): bg[b,p] = cls_map[b,p]/max < thresh (max = gmax[0], or the row max when
 * per_image); key_bias[b,0] = 0, key_bias[b,1+p] = -100*bg.  forced_bg [B,P] overrides the decision when non-NULL. */
VTC_API int vtc_cls_mask(const float* cls_map, const float* gmax, const uint8_t* forced_bg, float thresh, int32_t per_image,
                 uint8_t* bg, float* key_bias, int32_t batch, int32_t n_tokens, void* stream);
/* vtc_cls_stat + vtc_cls_mask in ONE launch (what the forward runs per layer >= mask_from): `ticket` = one zeroed uint32 per
 * call (left non-zero); batch must not exceed what is resident at once (else the two kernels are launched).  mask_operands
 * (optional, zeroed once by the caller, batch * vtc_attention_mask_operand_bytes(n_tokens) bytes): the additive mask of
 * vit_model.py:348-361 as the ready-made tensor-core operands of the fast attention kernel, K_aug[key] = bias / scale and
 * Q_aug[row] = [row not masked] (inv_scale = 1 / attention scale), which vtc_attention_masked then fetches with bulk copies
 * instead of rebuilding them from key_bias for every (head, query tile). */
VTC_API size_t vtc_attention_mask_operand_bytes(int32_t n_tokens);
VTC_API int vtc_cls_stat_mask(const float* cls_rows, float* cls_map, float* gmax, const uint8_t* forced_bg, float thresh, int32_t per_image,
                      uint8_t* bg, float* key_bias, uint32_t* ticket, void* mask_operands, float inv_scale, int32_t batch, int32_t heads,
                      int32_t n_tokens, void* stream);
/* vtc_attention (fast path: no full P) with the precomputed mask operands of vtc_cls_stat_mask; bit-identical to
 * vtc_attention(qkv, key_bias, ...) */
VTC_API int vtc_attention_masked(const void* qkv, const float* key_bias, const void* mask_operands, void* out, float* cls_rows, int32_t batch,
                         int32_t n_tokens, int32_t heads, float scale, void* stream);
/* high-weight-patch head + final norm + head (vit_model.py:374-422). tokens [B,N,D] fp32 (block-L output).  The top-16
 * runs on cls_map / max exactly like vit_model.py:372,377 (max = gmax[0], the batch-global one, or the image's own when
 * gmax is NULL): descending, ties to the smaller index, NaN ranked first (torch.topk). */
VTC_API int vtc_topk_heads(const vtc_model* m, const float* tokens, const float* cls_map, const float* gmax, const int32_t* forced_topk,
                   float* logits, float* hwp_logits, float* hwp_tokens, int32_t* topk_idx, int32_t batch, void* stream);

/* ---- CAM / rollout / pseudo-label post-processing ----------------------------------------------
 * attention rollout (predict.py:215-232): attn_mean [L,B,N,N] -> row [B,P] = (e0^T A_{L-1} ... A_0)[1:] with
 * A_l = (mean_l + I)/rowsum, evaluated as a reverse vector-matrix chain. */
VTC_API int vtc_rollout(const float* attn_mean, float* row, int32_t layers, int32_t batch, int32_t n_tokens, void* stream);
/* The same chain on bf16 "rollout operands", the form the forward keeps the head means in: per layer
 * bf16 [B,N,ldr], ldr = vtc_rollout_operand_ld(N) = N + 2 rounded up to 8; a row = N values | zero padding | the fp32 sum of
 * the N rounded values in its last four bytes.  16-byte aligned rows: the kernel streams row blocks with bulk copies. */
VTC_API int32_t vtc_rollout_operand_ld(int32_t n_tokens);
VTC_API int vtc_rollout_operand_from_mean(const float* attn_mean, void* operand, int32_t batch, int32_t n_tokens, void* stream);
VTC_API int vtc_rollout_operands(const void* operands, float* row, int32_t layers, int32_t batch, int32_t n_tokens, void* stream);
/* vtc_attention_mean writing the rollout operand of the layer (attn_mean and / or operand non-NULL) */
VTC_API int vtc_attention_mean_operand(const void* qkv, const float* key_bias, void* out, float* cls_rows, float* attn_mean, void* operand,
                               void* scratch, size_t scratch_bytes, int32_t batch, int32_t n_tokens, int32_t heads, float scale, void* stream);
/* per-layer CLS maps / bg map (predict.py:261-266, validate.py:225-237): cls_rows [L,B,H,N] ->
 * maps [B,P]: mean over layers [first,last) and heads, + identity on the CLS entry, / rowsum, patches, / max. */
VTC_API int vtc_cls_layer_map(const float* cls_rows, float* map, int32_t layers, int32_t first, int32_t last, int32_t batch,
                      int32_t heads, int32_t n_tokens, void* stream);
/* classic CAM (t.py:55-75, utils.py:80-88, vit_model.py:297): tokens [B,N,D] fp32, w [C,D] ->
 * cam [B,C,g,g] = minmax(relu(F.w^T)) per (b,c) map, eps added to the max. */
VTC_API int vtc_cam_project(const float* tokens, const float* w, float* cam, int32_t batch, int32_t n_tokens, int32_t dim,
                    int32_t classes, int32_t relu, float eps, void* stream);
/* row-wise divide by the row max: maps [rows,P] in place (predict.py:247,266) */
VTC_API int vtc_normalize_max(float* maps, int32_t rows, int32_t p, void* stream);
/* bilinear, align_corners=False (validate.py:177,239; cv2.resize predict.py:247): in [n,g,g] -> out [n,H,W] fp32 */
VTC_API int vtc_upsample_bilinear(const float* in, float* out, int32_t n, int32_t g, int32_t out_h, int32_t out_w, void* stream);
/* same, then *255 and truncate to uint8 (predict.py:266-269) */
VTC_API int vtc_upsample_bilinear_u8(const float* in, uint8_t* out, int32_t n, int32_t g, int32_t out_h, int32_t out_w,
                             void* stream);
/* CAM pseudo label (utils.py:100-108 convention): label[b,y,x] = argmax([bg_thresh, up(cam[b,c]) for labelled c]),
 * 0 = background, c+1 otherwise.  cam [B,C,g,g], labels [B,C] u8 -> out [B,H,W] u8.  Fused upsample + argmax. */
VTC_API int vtc_cam_label(const float* cam, const uint8_t* labels, float bg_thresh, uint8_t* out, int32_t batch,
                  int32_t classes, int32_t g, int32_t out_h, int32_t out_w, void* stream);
/* high-weight-patch class vote + cosine maps (validate.py:132-175): per image
 *   patch_to_cls[k] = mode{argmax_c W1'[c,d] : argmax_k ori[k,d] = k}, W1' rows of classes with sigmoid(hwp)<sig set to -10
 *   cos[k,p] = <ori[k]/|ori[k]|, F[p]/|F[p]|>
 * outputs patch_to_cls [B,K] int32 (class 0..C-1, or -1 when the patch owns no feature) and cos [B,K,g,g]. */
VTC_API int vtc_hwp_cos_vote(const float* hwp_logits, const float* head1_w, const float* hwp_tokens, const float* tokens,
                     float sig_thresh, int32_t* patch_to_cls, float* cos, int32_t batch, int32_t n_tokens, int32_t dim,
                     int32_t classes, int32_t k, void* stream);
/* validate.py:177-258 fused: seg[b,y,x] = (patch_to_cls[argmax_k up(cos[b,k])] + 1) * [max_k >= cos_thresh] *
 * [up(bg_map[b]) >= bg_thresh]; patches voting -1 count as background. out [B,H,W] u8. */
VTC_API int vtc_hwp_seg(const float* cos, const int32_t* patch_to_cls, const float* bg_map, float cos_thresh, float bg_thresh,
                uint8_t* out, int32_t batch, int32_t k, int32_t g, int32_t out_h, int32_t out_w, void* stream);
/* compute_mAP (utils.py:248-262): per-image sklearn average_precision_score over the C class scores, on the device.
 * labels [B,C] multi-hot fp32, scores [B,C] fp32 -> ap [B] fp64 (-1 for images without a positive label, which the
 * reference skips; may be NULL) and acc[0] += sum of APs, acc[1] += number of scored images (may be NULL). */
VTC_API int vtc_average_precision(const float* labels, const float* scores, int32_t batch, int32_t classes, double* ap, double* acc,
                          void* stream);
/* patch-token similarity of predict.py:191-199 (viz): F.normalize with its default dim=1 on [1,N,D] normalises every
 * feature column across the tokens; sim [B,N,N] is the gram matrix of the result.  scratch: [B,D] floats. */
VTC_API int vtc_patch_similarity(const float* tokens, float* scratch, float* sim, int32_t batch, int32_t n_tokens, int32_t dim,
                         void* stream);
/* ConfusionMatrix.update (utils.py:35-45): mat [n,n] int64 += bincount(n*gt + pred) over gt in [0,n). */
VTC_API int vtc_confmat_update(const uint8_t* gt, const uint8_t* pred, size_t count, int32_t n, int64_t* mat, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VTC_H_ */
