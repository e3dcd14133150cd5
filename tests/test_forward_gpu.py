"""End-to-end parity of the fused forward (through the drop-in VisionTransformer -> C-ABI -> sm_100a kernels) against
(a) the committed golden vectors produced by EXECUTING THE REFERENCE (tests/golden/make_golden.py) and
(b) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star, bf16 mode): logits max|d|/max|ref| <= 1e-2; CAM cosine >= 0.999; pseudo-label
argmax agreement >= 99.5 %.  Discrete decisions (background mask, top-16) are graded separately and the continuous
outputs are additionally compared with the decisions teacher-forced (SURVEY hard part 4)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

# BASELINE.json north_star, bf16 mode: logits max relative error <= 1e-2, CAM cosine >= 0.999, pseudo-label agreement >= 99.5 %.
# These bars are asserted (a) on the reference's own default-init weights (BASELINE config 2; the layer>4 mask never fires
# there) and (b) in the "masked" regime (oracle.vit_forward.masked: q/k rows x3.5 from block 3), where the background mask
# of vit_model.py:325-361 fires on 0.5-0.6 of the patches at a natural logit scale.
LOGIT_TOL = 1e-2
CAM_COS = 0.999
LABEL_AGREE = 0.995
# STRESS regime, reported separately and not a north-star claim: "peaked" = every qkv weight x5 (attention logits x25 AND the
# attention output x5 against the residual stream).  A CPU emulation of bf16-operand arithmetic (tools/emulate_bf16.py) sits
# at 0.9-1.2e-2 of the fp32 logits there, i.e. no bf16 implementation meets 1e-2 in it; the stress bound is 2e-2.
PEAKED_STRESS_TOL = 2e-2
PEAKED_TOL = PEAKED_STRESS_TOL


def relerr(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def cosine(a, b):
    a, b = torch.as_tensor(a).double().cpu().flatten(), torch.as_tensor(b).double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.fixture(scope="module")
def env(lib_built):
    import vision_transformer_cam_b200 as V
    from oracle import vit_forward as VF
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(dev).eval()
    model.is_train = False
    return dict(V=V, VF=VF, dev=dev, model=model, sd0=sd0, sd_peaked=VF.peaked(sd0), sd_masked=VF.masked(sd0))


def load(env, which):
    env["model"].load_state_dict({"default": env["sd0"], "peaked": env["sd_peaked"], "masked": env["sd_masked"]}[which]
                                 if isinstance(which, str) else which)
    return env["model"]


def test_default_regime_matches_reference_golden(env):
    gold = np.load(os.path.join(GOLDEN, "default_b2.npz"))
    model = load(env, "default")
    x = env["VF"].make_images(0, 2).to(env["dev"])
    assert abs(float(x.double().sum()) - float(gold["x_sig"][0])) < 1e-3
    o = model.forward_cam(x, tokens_layers=12, attn_layers=12, bg=True, cls_map=True)
    assert relerr(o.logits, gold["logits"]) <= LOGIT_TOL, relerr(o.logits, gold["logits"])
    assert relerr(o.hwp_logits, gold["hwp"]) <= 5e-2          # hwp also absorbs top-16 index flips in this flat regime
    assert relerr(o.tokens[:, :, 0, :], gold["x_cls"]) <= LOGIT_TOL
    assert relerr(o.tokens[-1][:, 7, :], gold["x_last_tok7"]) <= LOGIT_TOL
    # attention probabilities: near-uniform here (max 0.007); absolute tolerance 2 % of the max probability
    assert float((o.cls_rows.cpu() - torch.from_numpy(gold["cls_rows"])).abs().max()) <= 0.02 * float(gold["cls_rows"].max())
    assert float((o.attn[6][0, 3].cpu() - torch.from_numpy(gold["p_l6_img0_h3"])).abs().max()) <= 0.02 * float(gold["p_l6_img0_h3"].max())
    assert float((o.attn[-1][0].mean(0).cpu() - torch.from_numpy(gold["pbar_last_img0"])).abs().max()) <= 0.02 * float(gold["pbar_last_img0"].max())
    assert int(o.bg.sum()) == 0 and int(gold["bg"].sum()) == 0            # the mask never fires with the reference init
    for a, b in zip(o.tokens.abs().mean(dim=(1, 2, 3)).cpu().tolist(), gold["x_abs_mean"].tolist()):
        assert abs(a - b) <= 1e-2 * b


def test_reference_6tuple_surface(env):
    model = load(env, "default")
    x = env["VF"].make_images(0, 2).to(env["dev"])
    out = model(x)
    assert isinstance(out, tuple) and len(out) == 6
    logits, attn_w, attn_m, hwp, w1, ori = out
    assert logits.shape == (2, 20) and hwp.shape == (2, 20) and ori.shape == (2, 16, 768)
    assert len(attn_w) == 12 and attn_w[0].shape == (2, 12, 197, 197) and attn_w[0].dtype == torch.float32
    assert len(attn_m) == 12 and attn_m[0].shape == (2, 197, 768)
    assert w1.data_ptr() == model.head1.weight.data_ptr()                 # aliases the parameter like `.data`
    assert float((attn_w[3].sum(-1) - 1).abs().max()) < 1e-5
    with pytest.raises(AssertionError):
        model(torch.zeros(1, 3, 200, 200, device=env["dev"]))
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 224, 224))                                # CPU input: loud failure, no fallback


def test_masked_regime_meets_the_north_star_bar(env):
    """The reference's distinguishing feature -- the layer>4 background mask -- firing (bg fraction 0.52-0.58 on layers
    5..11), at the north-star tolerance: logits / hwp / tokens <= 1e-2 with the discrete decisions teacher-forced (so the
    continuous arithmetic is graded on identical masks), then the decisions themselves free-running."""
    gold = np.load(os.path.join(GOLDEN, "masked_b3.npz"))
    model = load(env, "masked")
    x = env["VF"].make_images(0, 3).to(env["dev"])
    frac = gold["bg"][1:].mean(axis=(1, 2))                               # masks consumed by layers 6..11 (+ the unused last)
    assert (frac > 0.1).all() and (frac < 0.9).all(), frac               # the mask path is really exercised
    forced = {4 + i: torch.from_numpy(gold["bg"][i]) for i in range(gold["bg"].shape[0])}
    o = model.forward_cam(x, tokens_layers=12, bg=True, cls_map=True, forced_bg=forced, forced_topk=torch.from_numpy(gold["topk_idx"]))
    e = dict(logits=relerr(o.logits, gold["logits"]), hwp=relerr(o.hwp_logits, gold["hwp"]), ori=relerr(o.hwp_tokens, gold["ori"]),
             x_cls=relerr(o.tokens[:, :, 0, :], gold["x_cls"]))
    print("masked teacher-forced:", {k: f"{v:.2e}" for k, v in e.items()})
    assert max(e.values()) <= LOGIT_TOL, e
    assert float((o.cls_rows.cpu() - torch.from_numpy(gold["cls_rows"])).abs().max()) <= 0.02 * float(gold["cls_rows"].max())
    assert torch.equal(o.bg[4:].cpu(), torch.from_numpy(gold["bg"]))
    assert cosine(o.cls_map[-1], gold["c_last"]) >= 0.9999
    # free running: the same forward making its own decisions
    f = model.forward_cam(x, bg=True, cls_map=True)
    agree = float((f.bg[4:].cpu() == torch.from_numpy(gold["bg"])).float().mean())
    overlap = np.mean([len(set(f.topk_idx[b].cpu().tolist()) & set(gold["topk_idx"][b].tolist())) / 16.0 for b in range(3)])
    ef = relerr(f.logits, gold["logits"])
    print(f"masked free-running: bg agreement {agree:.4f}, top-16 overlap {overlap:.3f}, logits relerr {ef:.2e}")
    assert agree >= 0.99 and overlap >= 0.9
    assert ef <= 3e-2                                                      # a flipped threshold-adjacent patch changes a whole key column


def test_masked_regime_cam_rollout_pseudo_labels(env):
    """CAM agreement (BASELINE metric) in the mask-firing regime against the reference's outputs + exec'd reference lines."""
    from vision_transformer_cam_b200 import cam as CAM
    gold = np.load(os.path.join(GOLDEN, "masked_b1.npz"))
    model = load(env, "masked")
    x = env["VF"].make_images(0, 1).to(env["dev"])
    forced = {4 + i: torch.from_numpy(gold["bg"][i]) for i in range(gold["bg"].shape[0])}
    o = model.forward_cam(x, attn_mean=True, forced_bg=forced, forced_topk=torch.from_numpy(gold["topk_idx"]))
    hw = (375, 500)
    cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
    assert cosine(cam, gold["classic_cam"]) >= CAM_COS, cosine(cam, gold["classic_cam"])
    for key_in, key_out in (("cam_labels_in", "cam_label"), ("cam_labels_sig_in", "cam_label_sig")):
        lab = CAM.cam_pseudo_label(cam, torch.from_numpy(gold[key_in]).to(env["dev"]), hw)
        agree = float((lab.cpu() == torch.from_numpy(gold[key_out])).float().mean())
        print(f"masked regime {key_out}: {int(gold[key_in].sum())} labels, agreement {agree:.5f}")
        assert agree >= LABEL_AGREE, (key_out, agree)
    assert cosine(CAM.rollout_row(o.attn_mean), gold["rollout_row"]) >= CAM_COS
    assert cosine(CAM.rollout_map(o.attn_mean, hw), gold["rollout_up"].astype(np.float32)) >= CAM_COS
    assert cosine(CAM.layer_maps(o.cls_rows)[:, 0], gold["layer_maps14"]) >= CAM_COS
    seg, p2c, bgm = CAM.hwp_pseudo_seg(o, model.head1.weight.data, hw, return_parts=True)
    assert cosine(bgm, gold["val_bg_map"]) >= CAM_COS
    ref_seg = torch.from_numpy(gold["val_seg"]).to(torch.uint8)
    ref_seg = torch.where(ref_seg > 21, torch.zeros_like(ref_seg), ref_seg)
    free = float((seg[0].cpu() == ref_seg).float().mean())
    print(f"masked regime validate pseudo-seg agreement {free:.5f}")
    assert free >= 0.98, free


@pytest.mark.parametrize("regime", ["default", "masked"])
def test_batch_256_against_the_oracle(env, regime):
    """BASELINE config 2 shape (B = 256, mask_norm='batch': the global max of vit_model.py:335 spans all 256 images) against
    the CPU oracle run live on the same inputs, itself pinned to the reference's own B = 256 outputs (golden).  Bars over ALL
    256 images: logits <= 1e-2, CAM cosine >= 0.999 (every image), CAM pseudo-label agreement >= 99.5 % with the image-level
    label rows of the reference's voc12/cls_labels.npy."""
    from vision_transformer_cam_b200 import cam as CAM
    from oracle import postproc as PP
    VF = env["VF"]
    gold = np.load(os.path.join(GOLDEN, regime + "_b256.npz"))
    sd = env["sd0"] if regime == "default" else VF.masked(env["sd0"], qk_scale=VF.MASKED_B256_QK_SCALE)
    x = VF.make_images(0, 256)
    assert abs(float(x.double().sum()) - float(gold["x_sig"][0])) < 1e-2
    ref = VF.forward(sd, x, VF.VIT_B16_224, keep_P=False)
    assert float((ref["logits"] - torch.from_numpy(gold["logits"])).abs().max()) <= 1e-6          # the oracle IS the reference here
    assert torch.equal(ref["topk_idx"], torch.from_numpy(gold["topk_idx"].astype(np.int64)))
    bg_ref = torch.stack([b for b in ref["bg"] if b is not None]).to(torch.uint8)                     # [8,256,196]
    assert np.array_equal(np.packbits(bg_ref.numpy(), axis=-1), gold["bg_packed"])
    model = load(env, sd)
    xd = x.to(env["dev"])
    free = model.forward_cam(xd, bg=True, mask_norm="batch")
    if regime == "default":
        assert int(bg_ref.sum()) == 0 and int(free.bg.sum()) == 0
        o = free
    else:
        frac = bg_ref[1:].float().mean(dim=(1, 2))
        assert bool((frac > 0.1).all()) and bool((frac < 0.9).all()), frac
        agree = float((free.bg[4:].cpu() == bg_ref).float().mean())
        print(f"B=256 masked free-running: bg agreement {agree:.5f}, logits relerr {relerr(free.logits, ref['logits']):.2e}")
        assert agree >= 0.99
        o = model.forward_cam(xd, bg=True, mask_norm="batch", forced_bg={4 + i: bg_ref[i] for i in range(8)}, forced_topk=ref["topk_idx"])
    e, eh = relerr(o.logits, ref["logits"]), relerr(o.hwp_logits, ref["hwp"])
    cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
    cam_ref = PP.classic_cam(ref["X"][-1], sd["head1.weight"])
    cos = torch.nn.functional.cosine_similarity(cam.cpu().double().flatten(1), cam_ref.double().flatten(1), dim=1)
    labels = torch.from_numpy(gold["labels"]).float()
    hw = (375, 500)
    lab = CAM.cam_pseudo_label(cam, labels.to(env["dev"]), hw).cpu()
    per_image = []
    for i in range(0, 256, 32):
        r = PP.cam_pseudo_label(cam_ref[i:i + 32], labels[i:i + 32], hw)
        per_image.append((lab[i:i + 32] == r).float().mean(dim=(1, 2)))
    per_image = torch.cat(per_image)
    print(f"B=256 {regime}: logits {e:.2e} hwp {eh:.2e}  CAM cosine min {float(cos.min()):.6f} mean {float(cos.mean()):.6f}  "
          f"label agreement mean {float(per_image.mean()):.5f} min {float(per_image.min()):.5f} ({float(labels.sum(1).mean()):.2f} labels / image)")
    assert e <= LOGIT_TOL, e
    if regime == "masked":
        assert eh <= LOGIT_TOL, eh                       # top-16 teacher-forced; default regime: near-uniform map, index flips
    assert float(cos.min()) >= CAM_COS
    assert float(per_image.mean()) >= LABEL_AGREE


def test_fused_rollout_output_matches_the_attn_mean_path_and_the_reference(env):
    """forward_cam(rollout=True): the rollout row computed inside the forward from bf16 rollout operands == the row obtained
    from the fp32 attn_mean output of the same forward (storage rounding only), in every configuration of requested outputs
    (packed-P path, full-P path, both), and against the reference golden at the CAM bar."""
    from vision_transformer_cam_b200 import cam as CAM
    gold = np.load(os.path.join(GOLDEN, "masked_b1.npz"))
    model = load(env, "masked")
    x = env["VF"].make_images(0, 1).to(env["dev"])
    forced = {4 + i: torch.from_numpy(gold["bg"][i]) for i in range(gold["bg"].shape[0])}
    kw = dict(forced_bg=forced, forced_topk=torch.from_numpy(gold["topk_idx"]))
    a = model.forward_cam(x, rollout=True, **kw)
    b = model.forward_cam(x, rollout=True, attn_mean=True, **kw)
    c = model.forward_cam(x, rollout=True, attn_layers=12, **kw)           # head means from the full fp32 P
    d = model.forward_cam(x, rollout=True, attn_layers=3, attn_mean=True, **kw)
    ref_row = CAM.rollout_row(b.attn_mean)
    assert torch.equal(a.rollout, b.rollout) and torch.equal(a.logits, b.logits)
    for o in (a, c, d):
        assert relerr(o.rollout, ref_row) < 2e-3, relerr(o.rollout, ref_row)
        assert cosine(o.rollout, gold["rollout_row"]) >= CAM_COS
    assert cosine(CAM.rollout_map(a.rollout, (375, 500)), gold["rollout_up"].astype(np.float32)) >= CAM_COS
    # batch of 5 with per-image normalisation == image by image
    xs = env["VF"].make_images(20, 5).to(env["dev"])
    full = model.forward_cam(xs, rollout=True, mask_norm="image")
    one = model.forward_cam(xs[3:4], rollout=True, mask_norm="image")
    assert torch.equal(full.rollout[3:4], one.rollout)


def test_peaked_regime_teacher_forced_continuous_parity(env):
    gold = np.load(os.path.join(GOLDEN, "peaked_b3.npz"))
    model = load(env, "peaked")
    x = env["VF"].make_images(0, 3).to(env["dev"])
    forced = {4 + i: torch.from_numpy(gold["bg"][i]) for i in range(gold["bg"].shape[0])}
    o = model.forward_cam(x, tokens_layers=12, bg=True, cls_map=True, forced_bg=forced,
                          forced_topk=torch.from_numpy(gold["topk_idx"]))
    print("peaked teacher-forced: logits", relerr(o.logits, gold["logits"]), "hwp", relerr(o.hwp_logits, gold["hwp"]))
    assert relerr(o.logits, gold["logits"]) <= PEAKED_TOL, relerr(o.logits, gold["logits"])
    assert relerr(o.hwp_logits, gold["hwp"]) <= PEAKED_TOL, relerr(o.hwp_logits, gold["hwp"])
    assert relerr(o.hwp_tokens, gold["ori"]) <= PEAKED_TOL
    assert relerr(o.tokens[:, :, 0, :], gold["x_cls"]) <= PEAKED_TOL
    assert float((o.cls_rows.cpu() - torch.from_numpy(gold["cls_rows"])).abs().max()) <= 0.02 * float(gold["cls_rows"].max())
    assert torch.equal(o.bg[4:].cpu(), torch.from_numpy(gold["bg"]))
    assert cosine(o.cls_map[-1], gold["c_last"]) >= 0.999


def test_peaked_regime_free_running_decisions(env):
    gold = np.load(os.path.join(GOLDEN, "peaked_b3.npz"))
    model = load(env, "peaked")
    x = env["VF"].make_images(0, 3).to(env["dev"])
    o = model.forward_cam(x, bg=True, cls_map=True)
    frac = gold["bg"].mean()
    assert 0.1 < frac < 0.9                                              # the mask path is really exercised
    agree = float((o.bg[4:].cpu() == torch.from_numpy(gold["bg"])).float().mean())
    overlap = np.mean([len(set(o.topk_idx[b].cpu().tolist()) & set(gold["topk_idx"][b].tolist())) / 16.0 for b in range(3)])
    print(f"bg agreement {agree:.4f}, top-16 overlap {overlap:.3f}, logits relerr {relerr(o.logits, gold['logits']):.4f}")
    assert agree >= 0.97                                                  # bf16 flips a few threshold-adjacent patches
    assert overlap >= 0.6
    assert relerr(o.logits, gold["logits"]) <= 0.15                       # flips cascade; bounded, not tight


def test_mask_norm_and_batch_independence(env):
    """mask_norm='image' makes every image independent of its batch: B=6 in one call == 3 calls of 2 (bit exact: the
    accumulation order of every kernel is fixed per row)."""
    model = load(env, "peaked")
    x = env["VF"].make_images(10, 6).to(env["dev"])
    full = model.forward_cam(x, mask_norm="image", bg=True)
    for i in range(0, 6, 2):
        part = model.forward_cam(x[i:i + 2], mask_norm="image", bg=True)
        assert torch.equal(part.logits, full.logits[i:i + 2])
        assert torch.equal(part.bg, full.bg[:, i:i + 2])
        assert torch.equal(part.tokens_last, full.tokens_last[i:i + 2])
    # batch-global max (reference semantics) couples the images: same call with mask_norm='batch' must still agree for
    # the image that holds the global max ... and differ from 'image' somewhere in the batch
    glob = model.forward_cam(x, mask_norm="batch", bg=True)
    assert not torch.equal(glob.bg, full.bg)


def test_cam_rollout_and_pseudo_labels_against_reference(env):
    """CAM agreement on the reference's predict/validate math (golden = reference outputs + exec'd reference lines)."""
    from vision_transformer_cam_b200 import cam as CAM
    gold = np.load(os.path.join(GOLDEN, "peaked_b1.npz"))
    model = load(env, "peaked")
    x = env["VF"].make_images(0, 1).to(env["dev"])
    forced = {4 + i: torch.from_numpy(gold["bg"][i]) for i in range(gold["bg"].shape[0])}
    o = model.forward_cam(x, attn_mean=True, forced_bg=forced, forced_topk=torch.from_numpy(gold["topk_idx"]))
    hw = (375, 500)
    # classic CAM (t.py:55-75 / utils.py:80-88)
    cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
    assert cosine(cam, gold["classic_cam"]) >= 0.999, cosine(cam, gold["classic_cam"])
    lab = CAM.cam_pseudo_label(cam, torch.from_numpy(gold["cam_labels_in"]).to(env["dev"]), hw)
    agree = float((lab.cpu() == torch.from_numpy(gold["cam_label"])).float().mean())
    assert agree >= 0.995, agree                                          # 2 image-level labels (a VOC cls_labels.npy row)
    lab8 = CAM.cam_pseudo_label(cam, torch.from_numpy(gold["cam_labels_sig_in"]).to(env["dev"]), hw)
    agree8 = float((lab8.cpu() == torch.from_numpy(gold["cam_label_sig"])).float().mean())
    print("cam label agreement: 2 labels", agree, " 8 labels (random-weight maps, near-tied argmax)", agree8)
    assert agree8 >= 0.9
    # attention rollout (predict.py:189-247) and per-layer CLS maps (predict.py:261-269)
    row = CAM.rollout_row(o.attn_mean)
    assert cosine(row, gold["rollout_row"]) >= 0.999
    up = CAM.rollout_map(o.attn_mean, hw)
    assert cosine(up, gold["rollout_up"].astype(np.float32)) >= 0.999
    lm = CAM.layer_maps(o.cls_rows)
    assert cosine(lm[:, 0], gold["layer_maps14"]) >= 0.999
    # validate.py:132-258 pseudo segmentation
    seg, p2c, bgm = CAM.hwp_pseudo_seg(o, model.head1.weight.data, hw, return_parts=True)
    assert cosine(bgm, gold["val_bg_map"]) >= 0.999
    ref_p2c = torch.from_numpy(gold["val_patch_to_cls"]).long()
    mine = p2c[0].cpu().long()
    vote_agree = float(((mine == ref_p2c) | ((mine < 0) & (ref_p2c >= 21))).float().mean())
    ref_seg = torch.from_numpy(gold["val_seg"]).to(torch.uint8)
    ref_seg = torch.where(ref_seg > 21, torch.zeros_like(ref_seg), ref_seg)
    free = float((seg[0].cpu() == ref_seg).float().mean())
    # the per-patch class is a mode over 768 argmax votes on random-weight features: a discrete decision that bf16 noise
    # flips for patches with near-tied votes (graded separately); with the reference's votes teacher-forced the label map
    # is compared at a 98 % bar (see below).  (Identical inputs give identical votes: tests/test_kernels_gpu.py.)
    from vision_transformer_cam_b200 import ops
    forced_p2c = torch.where(ref_p2c >= 21, torch.full_like(ref_p2c, -1), ref_p2c).to(torch.int32)[None].to(env["dev"])
    _, cos_maps = ops.hwp_cos_vote(o.hwp_logits, model.head1.weight.data, o.hwp_tokens, o.tokens_last.contiguous(), 0.9)
    seg_tf = ops.hwp_seg(cos_maps, forced_p2c, bgm, hw)
    forced_agree = float((seg_tf[0].cpu() == ref_seg).float().mean())
    print(f"validate pseudo-seg: vote agreement {vote_agree:.3f}, label agreement free {free:.4f}, votes forced {forced_agree:.4f}")
    assert vote_agree >= 0.75
    # 16-way argmax over cosine maps of random-weight tokens whose own bf16 deviation is ~1e-2 (PEAKED_TOL): near-tied
    # pixels flip; measured 0.985 -> bar 0.98 for this path, the 99.5 % bar holds for the CAM pseudo label above
    assert forced_agree >= 0.98, forced_agree
    assert free >= 0.8, free


def test_full_batch_256_properties(env):
    """BASELINE config 2 size (B=256): size-independent properties instead of a CPU oracle run.
    (1) P rows sum to 1 (via the CLS rows), (2) chunked == full under per-image normalisation, (3) outputs finite."""
    model = load(env, "default")
    g = torch.Generator(device=env["dev"]).manual_seed(1234)
    x = torch.randn((256, 3, 224, 224), generator=g, device=env["dev"])
    o = model.forward_cam(x, mask_norm="image")
    assert bool(torch.isfinite(o.logits).all()) and bool(torch.isfinite(o.tokens_last).all())
    assert float((o.cls_rows.sum(-1) - 1).abs().max()) < 1e-4
    part = model.forward_cam(x[64:96], mask_norm="image")
    assert torch.equal(part.logits, o.logits[64:96])
    assert torch.equal(part.hwp_tokens, o.hwp_tokens[64:96])


FP32_TOL = 1e-4           # BASELINE.json north_star: logits max relative error in fp32 mode


def test_fp32_mode_matches_reference_golden(env):
    """precision='fp32' (split-bf16 operands, 3 tensor-core products per contraction): logits within 1e-4 of the fp32
    reference on the reference's own weights, and on the peaked regime with the discrete decisions teacher-forced."""
    model = env["model"]
    model.set_precision("fp32")
    try:
        gold = np.load(os.path.join(GOLDEN, "default_b2.npz"))
        load(env, "default")
        x = env["VF"].make_images(0, 2).to(env["dev"])
        o = model.forward_cam(x, tokens_layers=12, bg=True)
        e, eh = relerr(o.logits, gold["logits"]), relerr(o.hwp_logits, gold["hwp"])
        et = relerr(o.tokens[:, :, 0, :], gold["x_cls"])
        print(f"fp32 mode default: logits {e:.2e} hwp {eh:.2e} cls tokens {et:.2e}")
        assert e <= FP32_TOL and et <= FP32_TOL
        assert float((o.cls_rows.cpu() - torch.from_numpy(gold["cls_rows"])).abs().max()) <= 1e-4 * float(gold["cls_rows"].max())
        assert torch.equal(o.topk_idx.cpu().long().sort(-1).values, torch.from_numpy(gold["topk_idx"]).long().sort(-1).values) or eh <= 1e-2
        gold = np.load(os.path.join(GOLDEN, "peaked_b3.npz"))
        load(env, "peaked")
        x = env["VF"].make_images(0, 3).to(env["dev"])
        forced = {4 + i: torch.from_numpy(gold["bg"][i]) for i in range(gold["bg"].shape[0])}
        o = model.forward_cam(x, tokens_layers=12, bg=True, forced_bg=forced, forced_topk=torch.from_numpy(gold["topk_idx"]))
        e, eh = relerr(o.logits, gold["logits"]), relerr(o.hwp_logits, gold["hwp"])
        print(f"fp32 mode peaked (teacher-forced): logits {e:.2e} hwp {eh:.2e}")
        assert e <= FP32_TOL and eh <= FP32_TOL and relerr(o.hwp_tokens, gold["ori"]) <= FP32_TOL
        # free running: with fp32-grade arithmetic the discrete decisions themselves agree
        o = model.forward_cam(x, bg=True)
        agree = float((o.bg[4:].cpu() == torch.from_numpy(gold["bg"])).float().mean())
        print(f"fp32 mode peaked free-running: bg agreement {agree:.5f} logits {relerr(o.logits, gold['logits']):.2e}")
        assert agree >= 0.999
    finally:
        model.set_precision("bf16")


def test_uint8_image_ingest_equals_fp32_forward(env):
    """forward_cam_u8 on decoded uint8 HWC images == forward_cam on Normalize(ToTensor(images)), bit for bit."""
    model = load(env, "peaked")
    g = torch.Generator().manual_seed(7)
    img = torch.randint(0, 256, (3, 224, 224, 3), generator=g, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    x = img.permute(0, 3, 1, 2).float().div(255)
    x = ((x - torch.tensor(mean)[None, :, None, None]) / torch.tensor(std)[None, :, None, None]).contiguous()
    a = model.forward_cam(x.to(env["dev"]), bg=True)
    b = model.forward_cam_u8(img.to(env["dev"]), mean, std, bg=True)
    assert torch.equal(a.logits, b.logits) and torch.equal(a.tokens, b.tokens) and torch.equal(a.bg, b.bg)
    with pytest.raises(ValueError):
        model.forward_cam_u8(x.to(env["dev"]))


LN_FUSED_WORKER = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import vision_transformer_cam_b200 as V
from oracle import vit_forward as VF
gold = np.load(os.path.join(sys.argv[1], "tests", "golden", "default_b2.npz"))
torch.manual_seed(0)
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to("cuda:0").eval()
o = model.forward_cam(VF.make_images(0, 2).to("cuda:0"), tokens_layers=12, attn_mean=True, attn_layers=12)
ref = torch.from_numpy(gold["logits"]).double()
# the head mean requested next to the LayerNorm-fused GEMMs comes from the packed-P attention variant: it must equal the
# mean of the full P the same forward returns (it used to be left unwritten on this path)
for l in range(12):
    d = float((o.attn_mean[l] - o.attn[l].mean(dim=1)).abs().max())
    assert d <= 2e-3 * float(o.attn[l].max()), (l, d)
o2 = model.forward_cam(VF.make_images(0, 2).to("cuda:0"), attn_mean=True)
pb = torch.from_numpy(gold["pbar_last_img0"])
assert float((o2.attn_mean[-1][0].cpu() - pb).abs().max()) <= 0.02 * float(pb.max())
e = float((o.logits.double().cpu() - ref).abs().max() / ref.abs().max())
et = float((o.tokens[:, :, 0, :].double().cpu() - torch.from_numpy(gold["x_cls"]).double()).abs().max() / np.abs(gold["x_cls"]).max())
print("LNFUSED logits %.3e tokens %.3e" % (e, et))
assert e <= 1e-2 and et <= 1e-2
'''


def test_ln_fused_forward_matches_reference_golden(tmp_path):
    """VTC_LN_FUSION=1: the forward without LayerNorm kernels (LayerNorm folded into the GEMMs, gemm.cu) meets the same bf16
    bar against the reference golden.  The switch is read once per process, hence the subprocess."""
    import subprocess
    import sys
    from conftest import ROOT
    script = tmp_path / "w.py"
    script.write_text(LN_FUSED_WORKER)
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600, env={**os.environ, "VTC_LN_FUSION": "1"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "LNFUSED" in r.stdout


def test_cuda_graph_replay_equals_eager_and_follows_weight_updates(env):
    """forward_cam_graphed (the launch-bound batch-1 path of predict.py / validate.py): bit-identical to the eager forward,
    for new inputs, after a weight update (packed copies refreshed before the replay), for uint8 input, and per signature."""
    model = load(env, "peaked")
    dev = env["dev"]
    for B in (1, 3):
        for seed in (0, 1, 2):
            x = env["VF"].make_images(50 + seed, B).to(dev)
            g = model.forward_cam_graphed(x, attn_mean=True)
            got = {k: getattr(g, k).clone() for k in ("logits", "hwp_logits", "hwp_tokens", "tokens", "cls_rows", "attn_mean")}
            e = model.forward_cam(x, attn_mean=True)
            for k, v in got.items():
                assert torch.equal(v, getattr(e, k)), (B, seed, k)
    assert len(model._engine.graphs) == 2
    # in-place weight update (what an optimiser step or load_state_dict does): same storage, new version
    x = env["VF"].make_images(60, 1).to(dev)
    before = model.forward_cam_graphed(x, attn_mean=True).logits.clone()
    model = load(env, "default")
    after = model.forward_cam_graphed(x, attn_mean=True).logits.clone()
    assert not torch.equal(before, after) and torch.equal(after, model.forward_cam(x).logits)
    assert len(model._engine.graphs) == 2                       # replayed, not re-captured
    # uint8 ingest through the same mechanism
    u8 = torch.randint(0, 256, (1, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).to(dev)
    assert torch.equal(model.forward_cam_graphed(u8).logits, model.forward_cam_u8(u8).logits)
    with pytest.raises(ValueError):
        model.forward_cam_graphed(x, forced_topk=torch.zeros((1, 16), dtype=torch.int32))
    # precision round trip: fp32 mode needs a larger packed-weight buffer, the bf16 graph captured above pointed into the
    # old one -- it must be re-captured, not replayed against freed memory
    want = model.forward_cam(x).logits.clone()
    assert torch.equal(model.forward_cam_graphed(x).logits, want)
    model.set_precision("fp32")
    try:
        f32 = model.forward_cam_graphed(x).logits.clone()
        assert torch.equal(f32, model.forward_cam(x).logits)
    finally:
        model.set_precision("bf16")
    junk = [torch.full((1 << 26,), 7, dtype=torch.uint8, device=dev) for _ in range(8)]      # recycle freed blocks
    assert torch.equal(model.forward_cam_graphed(x).logits, want)
    del junk
