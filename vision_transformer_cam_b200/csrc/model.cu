// Model handle + the fused forward: VisionTransformer.forward of the reference (vit_model.py:303-424) as a fixed
// sequence of kernel launches on one stream.  Host-only state; no device allocation, no synchronisation.
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"
#include "ops.h"

struct vtc_model {
    vtc_config cfg;
    int N, P, D, H, L, C, HID, KP, R;
    int HD = 64;                 // head dimension; != 64 routes attention through the general-shape kernel (attention_generic.cu)
    bool generic_patch = false;  // patch size not a multiple of 8: element-wise patch matrix, K padded to a multiple of 64
    vtc_weights w;
    std::vector<vtc_layer_weights> lw;
    // packed bf16 GEMM weights (inside the caller's packed buffer)
    const __nv_bfloat16* patch_w = nullptr;
    struct LayerPacked {
        const __nv_bfloat16 *qkv, *proj, *fc1, *fc2;
        // bf16 mode: qkv / fc1 hold gamma * W (LayerNorm folded in), with the two epilogue vectors of the fold (gemm.cu)
        const float *qkv_g, *qkv_c, *fc1_g, *fc1_c;
    };
    std::vector<LayerPacked> lp;
    bool packed = false;
    bool split = false;          // fp32 mode: GEMM operands and activations as (hi | lo) bf16 halves
    bool packed_split = false;   // precision the packed buffer was written for
    // optional per-kernel timing (vtc_model_profile)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;     // pool, two events per span
    std::vector<int> prof_kind;           // kind of span i (events 2i, 2i+1)
};

namespace vtc {

static size_t seg(size_t elems, size_t elem_bytes) { return align_up(elems * elem_bytes, 256); }
constexpr int kRolloutLayers = 12;      // the reference keeps the last 12 layers' attention (vit_model.py:322); predict.py rolls those out

// VTC_LN_FUSION=1 (bf16 mode only) runs the forward without LayerNorm kernels: LayerNorm folded into the GEMMs either side
// of it (gemm.cu).  Measured on B200 at B = 256: 2 % faster over a 10-step burst (10.21 vs 10.44 ms), no gain once the run
// is long enough to sit at the 1 kW power cap (10.66 vs 10.65 ms over 20 steps: the LayerNorm kernels are low-power
// phases, and the GEMM epilogues that absorb them stretch the high-power ones).  The default therefore keeps the separate
// LayerNorm, whose bf16 rounding is also independent of the row mean.
static bool ln_fusion_enabled(const vtc_model* m) {
    static const bool on = []() { const char* e = getenv("VTC_LN_FUSION"); return e && e[0] == '1'; }();
    return !m->split && m->HD == 64 && on;
}

static size_t packed_bytes(const vtc_model* m) {
    const size_t D = m->D, HID = m->HID, eb = m->split ? 4 : 2;
    const size_t fold = ln_fusion_enabled(m) ? 2 * (seg(3 * D, 4) + seg(HID, 4)) : 0;      // g, c of the LayerNorm-folded qkv / fc1 GEMMs
    return seg(D * m->KP, eb) + m->L * (seg(3 * D * D, eb) + seg(D * D, eb) + 2 * seg(HID * D, eb) + fold);
}

// bf16 mode: the head mean of P comes from the packed bf16 P of the fast attention kernel (attention_mean, attention.cu)
// instead of a [B,H,N,N] fp32 round trip
static bool fused_mean_ok(const vtc_model* m) {
    static const bool off = []() { const char* e = getenv("VTC_NO_FUSED_MEAN"); return e && e[0] == '1'; }();
    return !off && !m->split && m->HD == 64 && m->N <= kAttentionFusedMeanMaxTokens;
}

struct Workspace {
    __nv_bfloat16 *hbuf, *patches, *y, *qkv, *ao;
    float *tok, *cls_rows, *cls_map, *key_bias, *gmax, *attn_tmp, *stats, *mean_tmp;
    unsigned int* ticket;        // cls_stat_mask: one arrival counter per layer (zeroed with gmax)
    uint8_t* aug;                // attention mask operands of every image (ops.h: AugLayout), bf16 fast path only
    size_t aug_bytes;
    void* mean_scratch;          // packed P of attention_mean
    size_t mean_scratch_bytes;
    __nv_bfloat16* rollops;      // rollout operands of the last min(L,12) layers [Lr,B,N,ldr] (ops.h)
    size_t bytes;
};

static Workspace carve(const vtc_model* m, int B, const vtc_outputs* o, uint8_t* base) {
    Workspace ws{};
    const size_t M = static_cast<size_t>(B) * m->N, D = m->D, N = m->N;
    size_t off = 0;
    auto take = [&](size_t elems, size_t eb) { uint8_t* p = base ? base + off : nullptr; off += seg(elems, eb); return p; };
    size_t hb = M * m->HID;
    const size_t pb = static_cast<size_t>(B) * m->P * m->KP;
    if (pb > hb) hb = pb;
    const size_t eb = m->split ? 4 : 2;      // bf16, or a (hi | lo) pair of bf16
    ws.hbuf = reinterpret_cast<__nv_bfloat16*>(take(hb, eb));
    ws.patches = ws.hbuf;   // the patch matrix is dead before the first fc1 writes hbuf
    ws.y = reinterpret_cast<__nv_bfloat16*>(take(M * D, eb));
    ws.qkv = reinterpret_cast<__nv_bfloat16*>(take(M * 3 * D, eb));
    ws.ao = reinterpret_cast<__nv_bfloat16*>(take(M * D, eb));
    const bool all_tokens = o && o->tokens && o->tokens_layers >= m->L;
    ws.tok = reinterpret_cast<float*>(take(M * D, 4));    // embedding output / running residual stream
    (void)all_tokens;
    ws.cls_rows = (o && o->cls_rows) ? nullptr : reinterpret_cast<float*>(take(static_cast<size_t>(B) * m->H * N, 4));
    ws.cls_map = (o && o->cls_map) ? nullptr : reinterpret_cast<float*>(take(static_cast<size_t>(B) * m->P, 4));
    ws.key_bias = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N, 4));
    ws.gmax = reinterpret_cast<float*>(take(2 * m->L, 4));
    ws.ticket = ws.gmax ? reinterpret_cast<unsigned int*>(ws.gmax + m->L) : nullptr;
    ws.stats = reinterpret_cast<float*>(take(M * (D / 128) * 2, 4));       // LayerNorm row statistics (bf16 mode)
    ws.aug_bytes = (!m->split && m->HD == 64 && m->L > m->cfg.mask_from + 1) ? static_cast<size_t>(B) * attention_aug_layout(m->N).per_image : 0;
    ws.aug = ws.aug_bytes ? take(ws.aug_bytes, 1) : nullptr;
    const bool want_mean = o && (o->attn_mean || o->rollout);
    const bool need_mean = want_mean && !(o->attn && o->attn_layers >= m->L);       // some layer's mean is not a by-product of its full P
    ws.attn_tmp = (need_mean && !fused_mean_ok(m)) ? reinterpret_cast<float*>(take(static_cast<size_t>(B) * m->H * N * N, 4)) : nullptr;
    ws.mean_scratch_bytes = (need_mean && fused_mean_ok(m)) ? attention_mean_scratch_bytes(B, m->N, m->H) : 0;
    ws.mean_scratch = ws.mean_scratch_bytes ? take(ws.mean_scratch_bytes, 1) : nullptr;
    // rollout output: operands of the layers the reference's rollout sees (the last 12, vit_model.py:322), and an fp32 [B,N,N]
    // staging matrix for layers whose head mean comes from a full fp32 P while attn_mean itself is not an output
    const int Lr = m->L < kRolloutLayers ? m->L : kRolloutLayers;
    ws.rollops = (o && o->rollout) ? reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(Lr) * B * N * rollout_operand_ld(m->N), 2)) : nullptr;
    ws.mean_tmp = (o && o->rollout && !o->attn_mean && (!fused_mean_ok(m) || (o->attn && o->attn_layers > 0)))
                      ? reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * N, 4)) : nullptr;
    ws.bytes = off;
    return ws;
}

struct Span {
    vtc_model* m;
    cudaStream_t st;
    bool on;
    Span(vtc_model* m_, cudaStream_t st_, int kind) : m(m_), st(st_), on(m_->prof_on) {
        if (!on) return;
        const size_t i = m->prof_kind.size();
        while (m->prof_ev.size() < 2 * (i + 1)) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
            m->prof_ev.push_back(e);
        }
        m->prof_kind.push_back(kind);
        cudaEventRecord(m->prof_ev[2 * i], st);
    }
    ~Span() {
        if (on) cudaEventRecord(m->prof_ev[2 * (m->prof_kind.size() - 1) + 1], st);
    }
};
#define VTC_STEP(kind, call)                       \
    do {                                           \
        Span _span(m, st, kind);                   \
        if ((rc = (call)) != VTC_OK) return rc;    \
    } while (0)

// Image input of the forward: normalised fp32 NCHW (the reference's tensor) or decoded uint8 HWC + Normalize constants.
struct ImageInput {
    const float* f32 = nullptr;
    const uint8_t* u8 = nullptr;
    const float* mean = nullptr;   // host, 3 floats
    const float* std = nullptr;    // host, 3 floats
};

static int forward(vtc_model* m, const ImageInput& in, int B, const vtc_outputs* o, const vtc_forcing* f, void* workspace, size_t ws_bytes,
                   uint32_t flags, cudaStream_t st) {
    VTC_REQUIRE(m && (in.f32 || (in.u8 && in.mean && in.std)) && o && workspace, VTC_ERR_ARG, "forward: null pointer");
    VTC_REQUIRE(!in.u8 || m->cfg.in_c == 3, VTC_ERR_SHAPE, "forward: uint8 HWC input needs in_c == 3");
    VTC_REQUIRE(!m->split || (m->HD == 64 && !m->generic_patch), VTC_ERR_SHAPE,
                "forward: the fp32 (split) mode needs head_dim 64 and a patch size that is a multiple of 8");
    VTC_REQUIRE(m->packed, VTC_ERR_ARG, "forward: vtc_model_pack_weights has not been called");
    VTC_REQUIRE(B > 0, VTC_ERR_SHAPE, "forward: batch %d", B);
    VTC_REQUIRE(o->logits && o->hwp_logits && o->hwp_tokens, VTC_ERR_ARG, "forward: logits / hwp_logits / hwp_tokens are required outputs");
    VTC_REQUIRE(((flags & VTC_FWD_FP32_SPLIT) != 0) == m->split, VTC_ERR_ARG,
                "forward: VTC_FWD_FP32_SPLIT must match the model's precision (vtc_model_set_precision)");
    VTC_REQUIRE(m->packed_split == m->split, VTC_ERR_ARG, "forward: the weights were packed for the other precision; pack them again");
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VTC_ERR_WORKSPACE, "forward: workspace must be 256-byte aligned");
    int rc = check_arch();
    if (rc != VTC_OK) return rc;
    Workspace ws = carve(m, B, o, static_cast<uint8_t*>(workspace));
    VTC_REQUIRE(ws.bytes <= ws_bytes, VTC_ERR_WORKSPACE, "forward: workspace %zu bytes < required %zu", ws_bytes, ws.bytes);

    const int N = m->N, D = m->D, H = m->H, L = m->L, P = m->P, HID = m->HID;
    const int M = B * N;
    const size_t tok_elems = static_cast<size_t>(M) * D;
    const int Lr = L < kRolloutLayers ? L : kRolloutLayers;
    const int Lt = o->tokens ? o->tokens_layers : 0;
    const int La = o->attn ? o->attn_layers : 0;
    VTC_REQUIRE(Lt >= 0 && Lt <= L && La >= 0 && La <= L, VTC_ERR_ARG, "forward: tokens_layers / attn_layers out of range");
    VTC_REQUIRE(!o->tokens || Lt >= 1, VTC_ERR_ARG, "forward: tokens requested with tokens_layers == 0");
    const float scale = 1.0f / sqrtf(static_cast<float>(D / H));          // vit_model.py:97
    const int per_image = (flags & VTC_FWD_MASK_NORM_IMAGE) ? 1 : 0;
    const int sp = m->split ? 1 : 0;

    // ---- patch embedding + token assembly (vit_model.py:306-314)
    float* t_cur = ws.tok;
    if (in.u8 && m->generic_patch)
        VTC_STEP(VTC_PROF_PATCHIFY, patchify_generic_u8(in.u8, in.mean, in.std, ws.patches, B, m->cfg.img_size, m->cfg.patch_size, m->KP, st));
    else if (in.u8) VTC_STEP(VTC_PROF_PATCHIFY, patchify_u8(in.u8, in.mean, in.std, ws.patches, B, m->cfg.img_size, m->cfg.patch_size, st, sp));
    else if (m->generic_patch) VTC_STEP(VTC_PROF_PATCHIFY, patchify_generic(in.f32, ws.patches, B, m->cfg.in_c, m->cfg.img_size, m->cfg.patch_size, m->KP, st));
    else VTC_STEP(VTC_PROF_PATCHIFY, patchify(in.f32, ws.patches, B, m->cfg.in_c, m->cfg.img_size, m->cfg.patch_size, st, sp));
    VTC_STEP(VTC_PROF_PATCHIFY, cls_token_rows(m->w.cls_token, m->w.pos_embed, t_cur, B, N, D, st));
    VTC_STEP(VTC_PROF_GEMM_PATCH, gemm_bf16(ws.patches, m->patch_w, m->w.patch_b, nullptr, m->w.pos_embed, t_cur, B * P, D, m->KP, VTC_EPI_PATCH_EMBED, N, st, sp));
    VTC_CUDA(cudaMemsetAsync(ws.gmax, 0, sizeof(float) * 2 * L, st));      // + the per-layer tickets of cls_stat_mask
    // mask operands: written per layer by cls_stat_mask (one launch, every block resident), fetched by the attention producer
    const bool use_aug = ws.aug != nullptr && B <= cls_stat_mask_capacity() && N - 1 <= 2048;
    if (use_aug) VTC_CUDA(cudaMemsetAsync(ws.aug, 0, ws.aug_bytes, st));
    const void* aug = use_aug ? ws.aug : nullptr;
    if (o->bg) VTC_CUDA(cudaMemsetAsync(o->bg, 0, static_cast<size_t>(L) * B * P, st));

    bool have_bias = false;
    float* last_map = nullptr;
    // bf16 mode: no LayerNorm kernels, see gemm.cu (the split-precision mode keeps the separate LayerNorm)
    const bool ln_fused = ln_fusion_enabled(m);
    const bool fused_mean = fused_mean_ok(m);
    static const bool fuse_proj = []() { const char* e = getenv("VTC_LN_FUSE_PROJ"); return !(e && e[0] == '0'); }();
    if (ln_fused) VTC_STEP(VTC_PROF_LAYERNORM, residual_prep(t_cur, ws.y, ws.stats, M, D, st));
    // Alternating sweep direction: every kernel of the chain walks its rows / images in the opposite order of its
    // predecessor, so it starts on the data the predecessor wrote last, which is still in the 126 MB L2 (the activations
    // of one layer, 77-310 MB each at B = 256, do not fit as a whole: a same-direction sweep would miss on everything).
    static const bool serpentine = []() { const char* e = getenv("VTC_NO_SERPENTINE"); return !(e && e[0] == '1'); }();
    int dir = 0;
    auto next_dir = [&]() { const int d = dir; if (serpentine) dir ^= 1; return d; };
    for (int l = 0; l < L; ++l) {
        const vtc_layer_weights& w = m->lw[l];
        const vtc_model::LayerPacked& pw = m->lp[l];
        float* t_in = t_cur;
        float* t_out = t_cur;
        if (o->tokens && l >= L - Lt) t_out = o->tokens + static_cast<size_t>(l - (L - Lt)) * tok_elems;
        const bool want_map = (l >= m->cfg.mask_from) || (l == L - 1) || (o->cls_map != nullptr);
        float* cls_l = o->cls_rows ? o->cls_rows + static_cast<size_t>(l) * B * H * N : (want_map ? ws.cls_rows : nullptr);
        float* attn_l = nullptr;
        float* mean_l = o->attn_mean ? o->attn_mean + static_cast<size_t>(l) * B * N * N : nullptr;
        __nv_bfloat16* op_l = (o->rollout && l >= L - Lr) ? ws.rollops + static_cast<size_t>(l - (L - Lr)) * B * N * rollout_operand_ld(N) : nullptr;
        bool packed_mean = false;         // head mean / rollout operand via the packed P of the fast attention kernel (bf16 mode)
        if (o->attn && l >= L - La) attn_l = o->attn + static_cast<size_t>(l - (L - La)) * B * H * N * N;
        else if ((mean_l || op_l) && fused_mean) packed_mean = true;
        else if (mean_l || op_l) attn_l = ws.attn_tmp;

        if (ln_fused) {
            // LayerNorm lives inside the GEMMs: ws.y = bf16(residual stream), ws.stats = its row statistics (gemm.cu)
            VTC_STEP(VTC_PROF_GEMM_QKV, gemm_lnfold(ws.y, pw.qkv, pw.qkv_c, pw.qkv_g, ws.stats, m->cfg.ln_eps, ws.qkv, M, 3 * D, D, 0, st, next_dir()));
            const float* kb = (l > m->cfg.mask_from && have_bias) ? ws.key_bias : nullptr;                   // vit_model.py:118
            if (packed_mean) VTC_STEP(VTC_PROF_ATTENTION, attention_mean_operand(ws.qkv, kb, ws.ao, cls_l, mean_l, op_l, ws.mean_scratch, ws.mean_scratch_bytes, B, N, H, scale, st, next_dir(), aug));
            else VTC_STEP(VTC_PROF_ATTENTION, attention(ws.qkv, kb, ws.ao, cls_l, attn_l, B, N, H, scale, st, next_dir(), aug));
            if (fuse_proj) {
                VTC_STEP(VTC_PROF_GEMM_PROJ, gemm_resid_ln(ws.ao, pw.proj, w.proj_b, t_in, t_out, ws.y, ws.stats, M, D, D, st, next_dir()));
            } else {      // A/B: proj through the L2 reduction + a row pass that produces bf16(t) and its statistics
                VTC_STEP(VTC_PROF_GEMM_PROJ, gemm_bf16(ws.ao, pw.proj, w.proj_b, t_in, nullptr, t_out, M, D, D, VTC_EPI_BIAS_RESIDUAL, 0, st, 0, next_dir()));
                VTC_STEP(VTC_PROF_LAYERNORM, residual_prep(t_out, ws.y, ws.stats, M, D, st));
            }
            VTC_STEP(VTC_PROF_GEMM_FC1, gemm_lnfold(ws.y, pw.fc1, pw.fc1_c, pw.fc1_g, ws.stats, m->cfg.ln_eps, ws.hbuf, M, HID, D, 1, st, next_dir()));
            VTC_STEP(VTC_PROF_GEMM_FC2, gemm_resid_ln(ws.hbuf, pw.fc2, w.fc2_b, t_out, t_out, ws.y, ws.stats, M, D, HID, st, next_dir()));
        } else {
            VTC_STEP(VTC_PROF_LAYERNORM, layernorm_bf16(t_in, w.norm1_w, w.norm1_b, ws.y, M, D, m->cfg.ln_eps, st, sp, next_dir()));
            VTC_STEP(VTC_PROF_GEMM_QKV, gemm_bf16(ws.y, pw.qkv, w.qkv_b, nullptr, nullptr, ws.qkv, M, 3 * D, D, VTC_EPI_BIAS, 0, st, sp, next_dir()));
            const float* kb = (l > m->cfg.mask_from && have_bias) ? ws.key_bias : nullptr;                       // vit_model.py:118
            if (m->HD != 64) VTC_STEP(VTC_PROF_ATTENTION, attention_generic(ws.qkv, kb, ws.ao, cls_l, attn_l, B, N, H, m->HD, scale, st));
            else if (sp) VTC_STEP(VTC_PROF_ATTENTION, attention_kv(ws.qkv, kb, ws.ao, cls_l, attn_l, B, N, H, scale, true, st, next_dir()));
            else if (packed_mean) VTC_STEP(VTC_PROF_ATTENTION, attention_mean_operand(ws.qkv, kb, ws.ao, cls_l, mean_l, op_l, ws.mean_scratch, ws.mean_scratch_bytes, B, N, H, scale, st, next_dir(), aug));
            else VTC_STEP(VTC_PROF_ATTENTION, attention(ws.qkv, kb, ws.ao, cls_l, attn_l, B, N, H, scale, st, next_dir(), aug));
            VTC_STEP(VTC_PROF_GEMM_PROJ, gemm_bf16(ws.ao, pw.proj, w.proj_b, t_in, nullptr, t_out, M, D, D, VTC_EPI_BIAS_RESIDUAL, 0, st, sp, next_dir()));
            VTC_STEP(VTC_PROF_LAYERNORM, layernorm_bf16(t_out, w.norm2_w, w.norm2_b, ws.y, M, D, m->cfg.ln_eps, st, sp, next_dir()));
            VTC_STEP(VTC_PROF_GEMM_FC1, gemm_bf16(ws.y, pw.fc1, w.fc1_b, nullptr, nullptr, ws.hbuf, M, HID, D, VTC_EPI_BIAS_GELU, 0, st, sp, next_dir()));
            VTC_STEP(VTC_PROF_GEMM_FC2, gemm_bf16(ws.hbuf, pw.fc2, w.fc2_b, t_out, nullptr, t_out, M, D, HID, VTC_EPI_BIAS_RESIDUAL, 0, st, sp, next_dir()));
        }
        t_cur = t_out;

        if ((mean_l || op_l) && !packed_mean) {          // head mean from the full fp32 P of this layer
            float* mdst = mean_l ? mean_l : ws.mean_tmp;
            VTC_STEP(VTC_PROF_HEAD_MEAN, head_mean(attn_l, mdst, B, H, N, st));
            if (op_l) VTC_STEP(VTC_PROF_HEAD_MEAN, rollout_operand_from_mean(mdst, op_l, B, N, st));
        }
        if (want_map) {
            float* map_l = o->cls_map ? o->cls_map + static_cast<size_t>(l) * B * P : ws.cls_map;
            last_map = map_l;
            if (l >= m->cfg.mask_from) {                                                                        // vit_model.py:325
                const uint8_t* forced = (f && f->bg && (f->bg_layer_mask >> l & 1u)) ? f->bg + static_cast<size_t>(l) * B * P : nullptr;
                uint8_t* bg_l = o->bg ? o->bg + static_cast<size_t>(l) * B * P : nullptr;
                VTC_STEP(VTC_PROF_CLS, cls_stat_mask(cls_l, map_l, ws.gmax + l, forced, m->cfg.mask_thresh, per_image, bg_l, ws.key_bias, ws.ticket + l, B, H, N, st, use_aug ? ws.aug : nullptr, 1.0f / scale));
                have_bias = true;
            } else {
                VTC_STEP(VTC_PROF_CLS, cls_stat(cls_l, map_l, ws.gmax + l, B, H, N, st));
            }
        }
    }
    HeadParams hp{m->w.norm_w, m->w.norm_b, m->w.pre_w, m->w.pre_b, m->w.head_w, m->w.head_b, m->w.head1_w, m->w.head1_b,
                  D, m->R, m->C, m->cfg.topk, N, m->cfg.ln_eps};
    if (o->rollout) VTC_STEP(VTC_PROF_ROLLOUT, rollout_operand(ws.rollops, o->rollout, Lr, B, N, st));      // predict.py:215-232
    VTC_STEP(VTC_PROF_HEADS, topk_heads(hp, t_cur, last_map, per_image ? nullptr : ws.gmax + (L - 1), f ? f->topk_idx : nullptr, o->logits, o->hwp_logits, o->hwp_tokens, o->topk_idx, B, st));
    return VTC_OK;
}

}  // namespace vtc

extern "C" {

int vtc_model_create(const vtc_config* cfg, vtc_model** out) {
    using namespace vtc;
    VTC_REQUIRE(cfg && out, VTC_ERR_ARG, "model_create: null pointer");
    VTC_REQUIRE(cfg->img_size > 0 && cfg->patch_size > 0 && cfg->img_size % cfg->patch_size == 0, VTC_ERR_SHAPE,
                "model_create: img_size %d / patch_size %d", cfg->img_size, cfg->patch_size);
    VTC_REQUIRE(cfg->embed_dim > 0 && cfg->num_heads > 0 && cfg->embed_dim % cfg->num_heads == 0, VTC_ERR_SHAPE,
                "model_create: embed_dim %d is not a multiple of num_heads %d", cfg->embed_dim, cfg->num_heads);
    const int hd = cfg->embed_dim / cfg->num_heads;
    // head_dim 64 runs on the tcgen05 attention kernels; other multiples of 16 up to 128 (ViT-H/14: 80) on the general-shape one
    VTC_REQUIRE(hd == 64 || (hd % 16 == 0 && hd >= 16 && hd <= 128), VTC_ERR_SHAPE, "model_create: head_dim %d (64, or a multiple of 16 up to 128)", hd);
    VTC_REQUIRE(cfg->embed_dim % 256 == 0 && cfg->mlp_hidden % 256 == 0, VTC_ERR_SHAPE,
                "model_create: embed_dim %d and mlp_hidden %d must be multiples of 256", cfg->embed_dim, cfg->mlp_hidden);
    VTC_REQUIRE(cfg->patch_size % 8 != 0 || (cfg->in_c * cfg->patch_size * cfg->patch_size) % 64 == 0, VTC_ERR_SHAPE,
                "model_create: in_c*patch^2 must be a multiple of 64");
    VTC_REQUIRE(cfg->depth > 0 && cfg->depth <= 32 && cfg->num_classes > 0, VTC_ERR_SHAPE, "model_create: depth %d classes %d", cfg->depth, cfg->num_classes);
    VTC_REQUIRE(cfg->representation_size == 0 || cfg->representation_size == cfg->embed_dim, VTC_ERR_SHAPE,
                "model_create: representation_size must be 0 or embed_dim (head1 consumes embed_dim features, vit_model.py:295,393)");
    const int g = cfg->img_size / cfg->patch_size;
    VTC_REQUIRE(cfg->topk > 0 && cfg->topk <= 64 && cfg->topk <= g * g, VTC_ERR_SHAPE, "model_create: topk %d", cfg->topk);
    VTC_REQUIRE(hd == 64 || g * g + 1 <= kAttentionGenericMaxTokens, VTC_ERR_SHAPE, "model_create: head_dim %d supports at most %d tokens (%d given)", hd,
                kAttentionGenericMaxTokens, g * g + 1);
    vtc_model* m = new (std::nothrow) vtc_model();
    VTC_REQUIRE(m != nullptr, VTC_ERR_ARG, "model_create: out of host memory");
    m->cfg = *cfg;
    m->P = g * g;
    m->N = m->P + 1;
    m->D = cfg->embed_dim;
    m->H = cfg->num_heads;
    m->L = cfg->depth;
    m->C = cfg->num_classes;
    m->HID = cfg->mlp_hidden;
    m->HD = hd;
    m->generic_patch = cfg->patch_size % 8 != 0;
    m->KP = (cfg->in_c * cfg->patch_size * cfg->patch_size + 63) / 64 * 64;      // == in_c*patch^2 unless generic_patch

    m->R = cfg->representation_size;
    *out = m;
    return VTC_OK;
}

int vtc_model_destroy(vtc_model* m) {
    if (m) for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
    delete m;
    return VTC_OK;
}

int vtc_model_profile(vtc_model* m, int32_t enable) {
    using namespace vtc;
    VTC_REQUIRE(m, VTC_ERR_ARG, "profile: null model");
    m->prof_on = enable != 0;
    m->prof_kind.clear();
    return VTC_OK;
}

int vtc_model_profile_read(vtc_model* m, float* ms_per_kind, int32_t* launches_per_kind) {
    using namespace vtc;
    VTC_REQUIRE(m && ms_per_kind && launches_per_kind, VTC_ERR_ARG, "profile_read: null pointer");
    for (size_t i = 0; i < m->prof_kind.size(); ++i) {
        VTC_CUDA(cudaEventSynchronize(m->prof_ev[2 * i + 1]));
        float ms = 0.f;
        VTC_CUDA(cudaEventElapsedTime(&ms, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]));
        ms_per_kind[m->prof_kind[i]] += ms;
        launches_per_kind[m->prof_kind[i]] += 1;
    }
    m->prof_kind.clear();
    return VTC_OK;
}

size_t vtc_model_packed_bytes(const vtc_model* m) { return m ? vtc::packed_bytes(m) : 0; }

int vtc_model_pack_weights(vtc_model* m, const vtc_weights* w, void* packed, size_t bytes, void* stream) {
    using namespace vtc;
    VTC_REQUIRE(m && w && packed, VTC_ERR_ARG, "pack_weights: null pointer");
    VTC_REQUIRE(w->num_layers == m->L && w->layers, VTC_ERR_ARG, "pack_weights: %d layers given, model has %d", w->num_layers, m->L);
    VTC_REQUIRE(bytes >= packed_bytes(m), VTC_ERR_WORKSPACE, "pack_weights: buffer %zu < %zu", bytes, packed_bytes(m));
    VTC_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0, VTC_ERR_WORKSPACE, "pack_weights: buffer must be 256-byte aligned");
    VTC_REQUIRE(w->cls_token && w->pos_embed && w->patch_w && w->patch_b && w->norm_w && w->norm_b && w->head_w && w->head_b &&
                    w->head1_w && w->head1_b, VTC_ERR_ARG, "pack_weights: missing top-level parameter");
    VTC_REQUIRE((m->R == 0) == (w->pre_w == nullptr), VTC_ERR_ARG, "pack_weights: pre_logits weights do not match representation_size");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    m->w = *w;
    m->lw.assign(w->layers, w->layers + w->num_layers);
    m->w.layers = m->lw.data();
    m->lp.resize(m->L);
    uint8_t* base = static_cast<uint8_t*>(packed);
    size_t off = 0;
    const size_t D = m->D, HID = m->HID;
    int rc;
    const bool split = m->split;
    auto pack = [&](const float* src, size_t rows, size_t cols, const __nv_bfloat16** dst) -> int {
        VTC_REQUIRE(src != nullptr, VTC_ERR_ARG, "pack_weights: missing GEMM weight");
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(base + off);
        off += seg(rows * cols, split ? 4 : 2);
        *dst = d;
        return split ? split_bf16(src, d, rows, cols, st) : cast_bf16(src, d, rows * cols, st);
    };
    if (m->generic_patch) {        // K = in_c * patch^2 padded with zero columns to KP
        VTC_REQUIRE(!split, VTC_ERR_SHAPE, "pack_weights: the fp32 (split) mode needs a patch size that is a multiple of 8");
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(base + off);
        off += seg(D * m->KP, 2);
        m->patch_w = d;
        if ((rc = cast_bf16_pad(w->patch_w, d, D, m->cfg.in_c * m->cfg.patch_size * m->cfg.patch_size, m->KP, st)) != VTC_OK) return rc;
    } else if ((rc = pack(w->patch_w, D, m->KP, &m->patch_w)) != VTC_OK) return rc;
    for (int l = 0; l < m->L; ++l) {
        const vtc_layer_weights& lw = m->lw[l];
        VTC_REQUIRE(lw.norm1_w && lw.norm1_b && lw.norm2_w && lw.norm2_b && lw.qkv_b && lw.proj_b && lw.fc1_b && lw.fc2_b, VTC_ERR_ARG,
                    "pack_weights: layer %d misses a vector parameter", l);
        if (!ln_fusion_enabled(m)) {
            if ((rc = pack(lw.qkv_w, 3 * D, D, &m->lp[l].qkv)) != VTC_OK) return rc;
            if ((rc = pack(lw.proj_w, D, D, &m->lp[l].proj)) != VTC_OK) return rc;
            if ((rc = pack(lw.fc1_w, HID, D, &m->lp[l].fc1)) != VTC_OK) return rc;
            if ((rc = pack(lw.fc2_w, D, HID, &m->lp[l].fc2)) != VTC_OK) return rc;
        } else {
            // qkv and fc1 sit behind a LayerNorm: gamma folded into the weights, beta / mean handled by two epilogue vectors
            auto fold = [&](const float* W, const float* gamma, const float* beta, const float* bias, size_t rows, const __nv_bfloat16** dst,
                            const float** g, const float** c) -> int {
                VTC_REQUIRE(W != nullptr, VTC_ERR_ARG, "pack_weights: missing GEMM weight");
                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(base + off);
                off += seg(rows * D, 2);
                float* gp = reinterpret_cast<float*>(base + off);
                off += seg(rows, 4);
                float* cp = reinterpret_cast<float*>(base + off);
                off += seg(rows, 4);
                *dst = d; *g = gp; *c = cp;
                return fold_ln(W, gamma, beta, bias, d, gp, cp, static_cast<int>(rows), static_cast<int>(D), st);
            };
            if ((rc = fold(lw.qkv_w, lw.norm1_w, lw.norm1_b, lw.qkv_b, 3 * D, &m->lp[l].qkv, &m->lp[l].qkv_g, &m->lp[l].qkv_c)) != VTC_OK) return rc;
            if ((rc = pack(lw.proj_w, D, D, &m->lp[l].proj)) != VTC_OK) return rc;
            if ((rc = fold(lw.fc1_w, lw.norm2_w, lw.norm2_b, lw.fc1_b, HID, &m->lp[l].fc1, &m->lp[l].fc1_g, &m->lp[l].fc1_c)) != VTC_OK) return rc;
            if ((rc = pack(lw.fc2_w, D, HID, &m->lp[l].fc2)) != VTC_OK) return rc;
        }
    }
    m->packed = true;
    m->packed_split = split;
    return VTC_OK;
}

int vtc_model_set_precision(vtc_model* m, int32_t precision) {
    using namespace vtc;
    VTC_REQUIRE(m, VTC_ERR_ARG, "set_precision: null model");
    VTC_REQUIRE(precision == VTC_PRECISION_BF16 || precision == VTC_PRECISION_FP32_SPLIT, VTC_ERR_ARG, "set_precision: unknown precision %d", precision);
    m->split = precision == VTC_PRECISION_FP32_SPLIT;
    return VTC_OK;
}

size_t vtc_workspace_bytes(const vtc_model* m, int32_t batch, const vtc_outputs* outs) {
    if (!m || batch <= 0) return 0;
    return vtc::carve(m, batch, outs, nullptr).bytes;
}

int vtc_forward(vtc_model* m, const float* x, int32_t batch, const vtc_outputs* outs, const vtc_forcing* forcing, void* workspace,
                size_t workspace_bytes, uint32_t flags, void* stream) {
    vtc::ImageInput in;
    in.f32 = x;
    return vtc::forward(m, in, batch, outs, forcing, workspace, workspace_bytes, flags, static_cast<cudaStream_t>(stream));
}

int vtc_forward_u8(vtc_model* m, const uint8_t* x, const float* mean, const float* std, int32_t batch, const vtc_outputs* outs,
                   const vtc_forcing* forcing, void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
    vtc::ImageInput in;
    in.u8 = x;
    in.mean = mean;
    in.std = std;
    return vtc::forward(m, in, batch, outs, forcing, workspace, workspace_bytes, flags, static_cast<cudaStream_t>(stream));
}

int vtc_topk_heads(const vtc_model* m, const float* tokens, const float* cls_map, const float* gmax, const int32_t* forced_topk, float* logits,
                   float* hwp_logits, float* hwp_tokens, int32_t* topk_idx, int32_t batch, void* stream) {
    using namespace vtc;
    VTC_REQUIRE(m && m->packed, VTC_ERR_ARG, "topk_heads: model without weights");
    HeadParams hp{m->w.norm_w, m->w.norm_b, m->w.pre_w, m->w.pre_b, m->w.head_w, m->w.head_b, m->w.head1_w, m->w.head1_b,
                  m->D, m->R, m->C, m->cfg.topk, m->N, m->cfg.ln_eps};
    return topk_heads(hp, tokens, cls_map, gmax, forced_topk, logits, hwp_logits, hwp_tokens, topk_idx, batch, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
