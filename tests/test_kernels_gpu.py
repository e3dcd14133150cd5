"""Per-kernel parity on the B200: each libvtc kernel (called through the C-ABI) against a plain fp32 torch expression
of the same op on the same seeded inputs.  Tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib_built):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from vision_transformer_cam_b200 import _lib
    _lib.load()
    _lib.check(_lib.load().vtc_check_device(), "vtc_check_device")
    return torch.device("cuda:0")


def _rand(shape, seed, dev, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 128), (300, 768, 768), (1000, 2304, 768),
                                   (197, 3072, 768), (394, 768, 3072), (50432, 768, 768)])
def test_gemm_bias(dev, M, N, K):
    from vision_transformer_cam_b200 import ops, _lib
    a = _rand((M, K), 1, dev).bfloat16()
    w = _rand((N, K), 2, dev, 0.05).bfloat16()
    bias = _rand((N,), 3, dev)
    out = ops.gemm_bf16(a, w, bias, _lib.EPI_BIAS)
    ref = a.float() @ w.float().t() + bias
    torch.cuda.synchronize()
    # bf16 output rounding (2^-9 relative) + fp32 accumulation order: 1e-2 of the output scale
    assert relerr(out.float(), ref) < 1e-2, relerr(out.float(), ref)
    assert float((out.float() - ref).abs().mean() / ref.abs().mean()) < 4e-3


def test_gemm_gelu(dev):
    from vision_transformer_cam_b200 import ops, _lib
    M, N, K = 777, 3072, 768
    a = _rand((M, K), 4, dev).bfloat16()
    w = _rand((N, K), 5, dev, 0.05).bfloat16()
    bias = _rand((N,), 6, dev)
    out = ops.gemm_bf16(a, w, bias, _lib.EPI_BIAS_GELU)
    ref = F.gelu(a.float() @ w.float().t() + bias)
    assert relerr(out.float(), ref) < 1e-2
    # the erf approximation itself (A&S 7.1.26) must be far below bf16 resolution: check where |x| is small
    assert float((out.float() - ref).abs().mean() / ref.abs().mean()) < 4e-3


def test_gemm_residual_inplace(dev):
    from vision_transformer_cam_b200 import ops, _lib
    M, N, K = 1000, 768, 3072
    a = _rand((M, K), 7, dev).bfloat16()
    w = _rand((N, K), 8, dev, 0.02).bfloat16()
    bias = _rand((N,), 9, dev)
    res = _rand((M, N), 10, dev)
    ref = res + a.float() @ w.float().t() + bias
    buf = res.clone()
    out = ops.gemm_bf16(a, w, bias, _lib.EPI_BIAS_RESIDUAL, residual=buf, out=buf)
    assert out.data_ptr() == buf.data_ptr()
    # fp32 output: only accumulation order differs -> 1e-5 of scale
    assert relerr(out, ref) < 2e-5, relerr(out, ref)
    out2 = ops.gemm_bf16(a, w, bias, _lib.EPI_BIAS_RESIDUAL, residual=res)
    assert relerr(out2, ref) < 2e-5


def test_gemm_patch_embed(dev):
    from vision_transformer_cam_b200 import ops, _lib
    B, P, D, K = 3, 196, 768, 768
    N = P + 1
    a = _rand((B * P, K), 11, dev).bfloat16()
    w = _rand((D, K), 12, dev, 0.05).bfloat16()
    bias = _rand((D,), 13, dev)
    pos = _rand((N, D), 14, dev)
    tokens = torch.full((B, N, D), 123.0, device=dev)
    ops.gemm_bf16(a, w, bias, _lib.EPI_PATCH_EMBED, pos=pos, out=tokens, tokens=N)
    ref = (a.float() @ w.float().t() + bias).view(B, P, D) + pos[1:]
    assert relerr(tokens[:, 1:], ref) < 2e-5
    assert bool((tokens[:, 0] == 123.0).all())          # CLS rows untouched


def test_gemm_rejects_bad_shapes(dev):
    from vision_transformer_cam_b200 import ops, _lib
    a = torch.zeros((128, 100), device=dev, dtype=torch.bfloat16)
    w = torch.zeros((256, 100), device=dev, dtype=torch.bfloat16)
    with pytest.raises(_lib.VtcError):
        ops.gemm_bf16(a, w, torch.zeros(256, device=dev))


# ---------------------------------------------------------------------------------------- elementwise
def test_cast_patchify_cls_rows(dev):
    from vision_transformer_cam_b200 import ops
    x = _rand((2, 3, 224, 224), 20, dev)
    p = ops.patchify(x, 16)
    ref = F.unfold(x, kernel_size=16, stride=16).transpose(1, 2).reshape(2 * 196, 768)   # k = c*256 + kh*16 + kw
    assert torch.equal(p, ref.bfloat16())
    v = _rand((1000003,), 21, dev)
    assert torch.equal(ops.cast_bf16(v), v.bfloat16())
    tokens = torch.zeros((2, 197, 768), device=dev)
    cls, pos = _rand((1, 1, 768), 22, dev), _rand((1, 197, 768), 23, dev)
    ops.cls_token_rows(cls, pos, tokens)
    assert torch.equal(tokens[:, 0], (cls[0, 0] + pos[0, 0]).expand(2, -1))
    assert float(tokens[:, 1:].abs().max()) == 0.0


@pytest.mark.parametrize("D,rows", [(768, 1), (768, 50432), (1024, 577 * 3), (256, 100)])
def test_layernorm(dev, D, rows):
    from vision_transformer_cam_b200 import ops
    x = _rand((rows, D), 30, dev, 3.0) + 0.5
    g, b = _rand((D,), 31, dev), _rand((D,), 32, dev)
    y = ops.layernorm_bf16(x, g, b, 1e-6)
    ref = F.layer_norm(x, (D,), g, b, 1e-6)
    # fp32 statistics; only the final bf16 rounding differs: 2^-8 relative per element
    assert float((y.float() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max()) + 1e-6
    assert float((y.float() - ref.bfloat16().float()).abs().mean()) < 1e-4


# ------------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, H, scale, kb):
    B, N, D3 = qkv.shape
    D = D3 // 3
    q, k, v = qkv.float().view(B, N, 3, H, D // H).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * scale
    if kb is not None:                       # reference mask: -100 * min(v_i + v_j, 1)  (vit_model.py:348-361)
        v_ = (kb != 0).float()
        s = s - 100.0 * torch.clamp(v_[:, :, None] + v_[:, None, :], max=1.0)[:, None]
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, D)
    return o, p


@pytest.mark.parametrize("B,N,H,masked", [(2, 197, 12, False), (3, 197, 12, True), (1, 50, 12, True), (2, 256, 4, False), (5, 129, 2, True)])
def test_attention(dev, B, N, H, masked):
    from vision_transformer_cam_b200 import ops
    qkv = _rand((B, N, 3 * H * 64), 40, dev, 1.5).bfloat16()
    kb = None
    if masked:
        g = torch.Generator().manual_seed(41)
        kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
        kb[:, 0] = 0
        kb = kb.to(dev)
    out, cls, attn = ops.attention(qkv, H, 0.125, key_bias=kb, want_cls=True, want_attn=True)
    ref_o, ref_p = _attn_ref(qkv, H, 0.125, kb)
    # P is fp32 softmax of an fp32-accumulated bf16 product: exp2/ex2.approx vs exp -> 2e-6 absolute on probabilities
    assert float((attn - ref_p).abs().max()) < 5e-6, float((attn - ref_p).abs().max())
    assert float((cls - ref_p[:, :, 0, :]).abs().max()) < 5e-6
    assert float((attn.sum(-1) - 1).abs().max()) < 1e-5
    # O: P rounded to bf16 before the second MMA + bf16 output: 1e-2 of scale
    assert relerr(out.float(), ref_o) < 1e-2, relerr(out.float(), ref_o)
    out2, cls2, _ = ops.attention(qkv, H, 0.125, key_bias=kb, want_cls=True, want_attn=False)
    # fast path = column-split single-pass kernel, full-P path = two-pass kernel: same math, different exponent offsets
    assert relerr(out2.float(), ref_o) < 1e-2 and float((cls2 - ref_p[:, :, 0, :]).abs().max()) < 5e-6
    hm = ops.head_mean(attn)
    assert float((hm - attn.mean(1)).abs().max()) < 1e-6


@pytest.mark.parametrize("amp", [4.0, 12.0])
def test_attention_large_dynamic_range(dev, amp):
    """Peaked logits: later key chunks exceed the first chunk's maximum by far more than 2^8, which exercises the lazy
    rescale of the single-pass softmax (P chunks already in TMEM, row sum and staged CLS row are rescaled)."""
    from vision_transformer_cam_b200 import ops
    B, N, H = 3, 197, 12
    qkv = _rand((B, N, 3 * H * 64), 45, dev, 1.0)
    qkv[:, :, : 2 * H * 64] *= amp                     # q and k
    qkv = qkv.bfloat16()
    g = torch.Generator().manual_seed(46)
    kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
    kb[:, 0] = 0
    kb = kb.to(dev)
    for bias in (None, kb):
        out, cls, _ = ops.attention(qkv, H, 0.125, key_bias=bias, want_cls=True, want_attn=False)      # fast (column-split) kernel
        out_ref, cls_ref, attn = ops.attention(qkv, H, 0.125, key_bias=bias, want_cls=True, want_attn=True)   # two-pass kernel
        ref_o, ref_p = _attn_ref(qkv, H, 0.125, bias)
        assert float((attn - ref_p).abs().max()) < 5e-5          # two-pass kernel, |logits| up to ~1e3 at amp 12
        assert float((cls - ref_p[:, :, 0, :]).abs().max()) < 2e-5, float((cls - ref_p[:, :, 0, :]).abs().max())
        assert relerr(out.float(), ref_o) < 1.5e-2, relerr(out.float(), ref_o)
        assert relerr(out.float(), out_ref.float()) < 1.5e-2


@pytest.mark.parametrize("B,N,H,masked", [(2, 577, 16, True), (2, 785, 12, True), (1, 785, 3, False), (3, 197, 12, True),
                                          (2, 256, 4, False), (5, 129, 2, True), (1, 50, 3, True), (1, 1025, 2, True)])
def test_attention_kv_blocked(dev, B, N, H, masked):
    """KV-blocked kernel (BASELINE configs 4 / 5: 785 and 577 tokens, and short sequences through the same code):
    O, CLS rows and the full P of the second sweep against fp32 torch."""
    from vision_transformer_cam_b200 import ops
    qkv = _rand((B, N, 3 * H * 64), 50, dev, 1.5).bfloat16()
    kb = None
    if masked:
        g = torch.Generator().manual_seed(51)
        kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
        kb[:, 0] = 0
        kb = kb.to(dev)
    ref_o, ref_p = _attn_ref(qkv, H, 0.125, kb)
    out, cls, attn = ops.attention_kv(qkv, H, 0.125, key_bias=kb, want_cls=True, want_attn=True)
    assert float((attn - ref_p).abs().max()) < 5e-6, float((attn - ref_p).abs().max())
    assert float((cls - ref_p[:, :, 0, :]).abs().max()) < 5e-6
    assert float((attn.sum(-1) - 1).abs().max()) < 2e-5
    assert relerr(out.float(), ref_o) < 1e-2, relerr(out.float(), ref_o)
    out2, cls2, _ = ops.attention_kv(qkv, H, 0.125, key_bias=kb, want_cls=True, want_attn=False)
    assert torch.equal(out2, out) and torch.equal(cls2, cls)
    # vtc_attention without the full-P request: the column-split pipelined kernel (attention_cs.cu), any N
    out3, cls3, _ = ops.attention(qkv, H, 0.125, key_bias=kb, want_cls=True, want_attn=False)
    assert relerr(out3.float(), ref_o) < 1e-2, relerr(out3.float(), ref_o)
    assert float((cls3 - ref_p[:, :, 0, :]).abs().max()) < 5e-6, float((cls3 - ref_p[:, :, 0, :]).abs().max())
    if N > 256:      # ... and with it: the KV-blocked kernel
        out4, cls4, attn4 = ops.attention(qkv, H, 0.125, key_bias=kb, want_cls=True, want_attn=True)
        assert torch.equal(out4, out) and torch.equal(cls4, cls) and torch.equal(attn4, attn)


@pytest.mark.parametrize("amp", [4.0, 12.0])
def test_attention_kv_large_dynamic_range(dev, amp):
    """Peaked logits over several key blocks: the lazy rescale must also rescale the O accumulator in TMEM."""
    from vision_transformer_cam_b200 import ops
    B, N, H = 2, 577, 4
    qkv = _rand((B, N, 3 * H * 64), 55, dev, 1.0)
    qkv[:, :, : 2 * H * 64] *= amp
    qkv = qkv.bfloat16()
    g = torch.Generator().manual_seed(56)
    kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
    kb[:, 0] = 0
    kb = kb.to(dev)
    for bias in (None, kb):
        out, cls, attn = ops.attention_kv(qkv, H, 0.125, key_bias=bias, want_cls=True, want_attn=True)
        ref_o, ref_p = _attn_ref(qkv, H, 0.125, bias)
        assert float((attn - ref_p).abs().max()) < 5e-5, float((attn - ref_p).abs().max())
        assert float((cls - ref_p[:, :, 0, :]).abs().max()) < 2e-5
        assert relerr(out.float(), ref_o) < 1.5e-2, relerr(out.float(), ref_o)
        out, cls, _ = ops.attention(qkv, H, 0.125, key_bias=bias, want_cls=True, want_attn=False)     # column-split kernel
        assert float((cls - ref_p[:, :, 0, :]).abs().max()) < 2e-5, float((cls - ref_p[:, :, 0, :]).abs().max())
        assert relerr(out.float(), ref_o) < 1.5e-2, relerr(out.float(), ref_o)


def test_attention_too_long_is_an_error(dev):
    from vision_transformer_cam_b200 import ops, _lib
    qkv = torch.zeros((1, 2049, 3 * 64), device=dev, dtype=torch.bfloat16)
    with pytest.raises(_lib.VtcError):
        ops.attention(qkv, 1, 0.125)


# -------------------------------------------------------------------------------------------- CLS ops
def test_cls_stat_and_mask(dev):
    from vision_transformer_cam_b200 import ops
    from oracle.vit_forward import cls_row_stat
    B, H, N = 5, 12, 197
    P = torch.rand((B, H, N, N), generator=torch.Generator().manual_seed(50)).softmax(-1)
    P[:, :, 0, 5:40] *= 8
    P = P / P.sum(-1, keepdim=True)
    ref = cls_row_stat(P)                                  # vit_model.py:329-334 on the full matrix
    cmap, gmax = ops.cls_stat(P[:, :, 0, :].contiguous().to(dev))
    assert float((cmap.cpu() - ref).abs().max()) < 1e-7
    assert abs(float(gmax) - float(ref.max())) < 1e-7
    bg, kb = ops.cls_mask(cmap, gmax, 0.25, per_image=False)
    ref_bg = torch.lt(ref / ref.max(), 0.25)
    assert float((bg.cpu().bool() != ref_bg).float().mean()) < 1e-3      # ties at the threshold only
    assert torch.equal(kb[:, 1:].cpu(), bg.cpu().float() * -100.0) and float(kb[:, 0].abs().max()) == 0
    bg_i, _ = ops.cls_mask(cmap, gmax, 0.25, per_image=True)
    ref_i = torch.lt(ref / ref.max(dim=1, keepdim=True).values, 0.25)
    assert float((bg_i.cpu().bool() != ref_i).float().mean()) < 1e-3
    forced = (torch.arange(B * (N - 1)).reshape(B, N - 1) % 3 == 0).to(torch.uint8).to(dev)
    bg_f, kb_f = ops.cls_mask(cmap, gmax, 0.25, forced_bg=forced)
    assert torch.equal(bg_f, forced)


@pytest.mark.parametrize("B,N,H", [(5, 197, 12), (3, 785, 12), (2, 577, 16), (4, 50, 12)])
def test_cls_stat_mask_one_launch_and_precomputed_mask_operands(dev, B, N, H):
    """The forward's one-launch mask builder == cls_stat + cls_mask, and the fast attention kernel fed with the mask operands it
    precomputes (bulk copies in the producer) is BIT-IDENTICAL to the kernel that rebuilds them from key_bias per item."""
    from vision_transformer_cam_b200 import ops
    g = torch.Generator().manual_seed(70 + N)
    cls_rows = torch.rand((B, H, N), generator=g).softmax(-1)
    cls_rows[:, :, 5:N // 3] *= 6
    cls_rows = (cls_rows / cls_rows.sum(-1, keepdim=True)).contiguous().to(dev)
    cmap0, gmax0 = ops.cls_stat(cls_rows)
    for per_image in (False, True):
        bg0, kb0 = ops.cls_mask(cmap0, gmax0, 0.25, per_image=per_image)
        cmap, gmax, bg, kb, aug = ops.cls_stat_mask(cls_rows, 0.25, per_image=per_image, scale=0.125)
        assert torch.equal(cmap, cmap0) and torch.equal(gmax, gmax0) and torch.equal(bg, bg0) and torch.equal(kb, kb0)
        assert 0.05 < float(bg.float().mean()) < 0.95
        qkv = (_rand((B, N, 3 * H * 64), 71, dev) * 1.5).bfloat16()
        o0, c0, _ = ops.attention(qkv, H, 0.125, key_bias=kb)
        o1, c1 = ops.attention_masked(qkv, H, 0.125, kb, aug)
        assert torch.equal(o0, o1) and torch.equal(c0, c1)
    forced = (torch.arange(B * (N - 1)).reshape(B, N - 1) % 3 == 0).to(torch.uint8).to(dev)
    _, _, bgf, kbf, augf = ops.cls_stat_mask(cls_rows, 0.25, forced_bg=forced, scale=0.125)
    assert torch.equal(bgf, forced)
    o0, c0, _ = ops.attention(qkv, H, 0.125, key_bias=kbf)
    o1, c1 = ops.attention_masked(qkv, H, 0.125, kbf, augf)
    assert torch.equal(o0, o1) and torch.equal(c0, c1)


def test_topk_heads_bit_exact(dev):
    """vit_model.py:372-393 is index work: on the reference's own CLS maps (golden c_last of the B = 256 runs, where the
    batch-global max of :372 spans 256 images) `vtc_topk_heads` must return torch.topk's indices IN ORDER, gather the
    block-L tokens exactly, and reproduce head1(mean) / head(LN(cls)) to fp32 rounding."""
    import os
    import numpy as np
    from conftest import GOLDEN
    import vision_transformer_cam_b200 as V
    torch.manual_seed(0)
    model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
    for name in ("default_b256", "masked_b256"):
        gold = np.load(os.path.join(GOLDEN, name + ".npz"))
        c = torch.from_numpy(gold["c_last"])                                   # [256,196] fp32, the reference's values
        B = c.shape[0]
        tokens = _rand((B, 197, 768), 77, dev)
        for per_image in (False, True):
            mx = c.max(dim=1, keepdim=True).values if per_image else c.max()
            m14 = c / mx                                                       # vit_model.py:372
            ref_idx = torch.stack([torch.topk(m14[j], 16, dim=0).indices for j in range(B)])       # :377
            if not per_image:
                assert torch.equal(ref_idx, torch.from_numpy(gold["topk_idx"].astype(np.int64)))   # == what the reference chose
            gmax = None if per_image else c.max().reshape(1).to(dev)
            logits, hwp, ori, idx = model.topk_heads(tokens, c.to(dev), gmax)
            idx = idx.cpu().long()
            # the selected VALUES, in order, are torch.topk's bit for bit; the INDICES are identical wherever the order is
            # defined.  Bit-equal fp32 quotients (one pair among these 256 x 16) have no defined order in torch.topk (its CPU
            # and CUDA kernels disagree with each other); libvtc returns the smaller index first.
            assert torch.equal(m14.gather(1, idx), m14.gather(1, ref_idx)), (name, per_image)
            v = m14.gather(1, ref_idx)
            tied = torch.zeros_like(v, dtype=torch.bool)
            tied[:, 1:] |= v[:, 1:] == v[:, :-1]
            tied[:, :-1] |= v[:, :-1] == v[:, 1:]
            assert torch.equal(idx[~tied], ref_idx[~tied]), (name, per_image)
            assert int(tied.sum()) <= 4 and bool((idx.sort(1).values == ref_idx.sort(1).values).all())     # same SET in every image
            for b, k in tied.nonzero().tolist():
                if k + 1 < 16 and v[b, k] == v[b, k + 1]:
                    assert idx[b, k] < idx[b, k + 1]
            ref_ori = torch.stack([tokens[j][idx[j].to(dev) + 1] for j in range(B)])               # :381-389
            assert torch.equal(ori, ref_ori)
            ref_hwp = F.linear(ref_ori.mean(dim=1), model.head1.weight, model.head1.bias)          # :392-393
            assert relerr(hwp, ref_hwp) <= 1e-6, relerr(hwp, ref_hwp)
            ref_logits = model.head(F.layer_norm(tokens[:, 0], (768,), model.norm.weight, model.norm.bias, 1e-6))
            assert relerr(logits, ref_logits) <= 1e-5, relerr(logits, ref_logits)
    # forced indices are used as given (and clamped into range instead of reading out of bounds)
    forced = torch.tensor([[0, 195, 7, 7] + list(range(12))] * B, dtype=torch.int32)
    _, _, ori_f, idx_f = model.topk_heads(tokens, c.to(dev), None, forced_topk=forced)
    assert torch.equal(idx_f.cpu(), forced) and torch.equal(ori_f[:, 1], tokens[:, 196])
    bad = forced.clone()
    bad[:, 0], bad[:, 1] = -5, 9999
    _, _, _, idx_b = model.topk_heads(tokens, c.to(dev), None, forced_topk=bad)
    assert int(idx_b.min()) == 0 and int(idx_b.max()) == 195


def test_topk_heads_ties_nan_and_inf(dev):
    """Edge cases of the selection: exact ties go to the smaller index, NaN ranks above every number (torch.topk), -inf
    entries stay selectable, and no index is returned twice (a NaN map used to index out of bounds)."""
    import vision_transformer_cam_b200 as V
    torch.manual_seed(0)
    model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
    P = 196
    maps = torch.zeros((6, P))
    maps[0] = torch.arange(P).float() % 7                       # many exact ties
    maps[1] = float("nan")                                      # all NaN
    maps[2] = torch.rand(P, generator=torch.Generator().manual_seed(1)); maps[2, [3, 50]] = float("nan")
    maps[3] = float("-inf"); maps[3, 10:20] = 1.0               # fewer than 16 finite entries
    maps[4] = torch.rand(P, generator=torch.Generator().manual_seed(2)); maps[4, 100] = float("inf")
    maps[5] = 0.5                                               # constant map
    tokens = _rand((6, 197, 768), 78, dev)
    one = torch.ones(1, device=dev)
    _, _, ori, idx = model.topk_heads(tokens, maps.to(dev), one)
    idx = idx.cpu().long()
    for b in range(6):
        assert len(set(idx[b].tolist())) == 16 and int(idx[b].min()) >= 0 and int(idx[b].max()) < P
        key = torch.where(torch.isnan(maps[b]), torch.full_like(maps[b], float("inf")), maps[b])
        order = sorted(range(P), key=lambda j: (-key[j].item(), j))[:16]          # descending value, ascending index
        assert idx[b].tolist() == order, (b, idx[b].tolist(), order)
        tv = torch.topk(maps[b], 16).values                                        # torch agrees on the VALUES (its tie order is unspecified)
        assert torch.equal(torch.nan_to_num(maps[b][idx[b]], nan=9e9), torch.nan_to_num(tv, nan=9e9))
    assert torch.equal(ori[0], tokens[0][idx[0].to(dev) + 1])


# ------------------------------------------------------------------------------------------- postproc
def test_rollout_and_layer_maps(dev):
    from vision_transformer_cam_b200 import ops
    from oracle import postproc as PP
    L, B, H, N = 12, 3, 12, 197
    g = torch.Generator().manual_seed(60)
    P_list = [(torch.randn((B, H, N, N), generator=g) * 2).softmax(-1) for _ in range(L)]
    mean = PP.head_mean(P_list)
    row = ops.rollout(mean.to(dev))
    ref = PP.rollout_dense(P_list)
    assert relerr(row.cpu(), ref) < 1e-5
    cls_rows = torch.stack([p[:, :, 0, :] for p in P_list]).contiguous().to(dev)
    bgm = ops.cls_layer_map(cls_rows, 5, L)
    assert float((bgm.cpu() - PP.bg_map(cls_rows.cpu())).abs().max()) < 1e-6
    lm = torch.stack([ops.cls_layer_map(cls_rows, l, l + 1) for l in range(L)])
    assert float((lm.cpu().view(L, B, 14, 14) - PP.layer_maps(P_list)).abs().max()) < 1e-6


@pytest.mark.parametrize("L,B,N", [(12, 3, 197), (12, 2, 785), (5, 4, 50), (1, 2, 197), (12, 1, 577)])
def test_rollout_operand_streaming_kernel(dev, L, B, N):
    """The forward's rollout path: head means as bf16 'rollout operands' (row = N values | zero padding | fp32 sum of the rounded
    values), streamed by bulk copies.  (1) Exact on the values it is given: equal to the fp32 dense chain of predict.py:221-232
    run on the SAME bf16-rounded matrices to 1e-5; (2) against the un-rounded fp32 chain the bf16 storage costs < 4e-3 relative
    and the cosine stays > 0.99999 (bar 0.999)."""
    from vision_transformer_cam_b200 import ops
    from oracle import postproc as PP
    g = torch.Generator().manual_seed(61 + N)
    mean = torch.stack([(torch.randn((B, N, N), generator=g) * 2).softmax(-1) for _ in range(L)])          # [L,B,N,N]
    opnd = ops.rollout_operand_from_mean(mean.to(dev))
    ldr = ops.rollout_operand_ld(N)
    assert opnd.shape == (L, B, N, ldr) and ldr % 8 == 0 and ldr >= N + 2
    rounded = mean.bfloat16().float()
    assert torch.equal(opnd[..., :N].float().cpu(), rounded) and float(opnd[..., N:ldr - 2].float().abs().max()) == 0.0
    rs = opnd[..., ldr - 2:].contiguous().view(torch.float32)[..., 0].cpu()
    assert float((rs - rounded.sum(-1)).abs().max()) < 1e-5
    row = ops.rollout_operands(opnd, N).cpu()

    def dense(m):                                              # predict.py:215-232 on head means m [L,B,N,N]
        aug = m.double() + torch.eye(N, dtype=torch.float64)
        aug = aug / aug.sum(-1, keepdim=True)
        joint = aug[0]
        for n in range(1, L):
            joint = aug[n] @ joint
        return joint[:, 0, 1:]

    assert relerr(row, dense(rounded)) < 1e-5, relerr(row, dense(rounded))
    full = dense(mean)
    assert relerr(row, full) < 4e-3          # one bf16 rounding (2^-9) at most; it averages out over the chain for L > 1
    cos = float((row.double().flatten() @ full.flatten()) / (row.double().norm() * full.norm()))
    assert cos > 0.99999, cos
    # and the fp32 entry point on the fp32 matrices
    assert relerr(ops.rollout(mean.to(dev)).cpu(), full) < 1e-5


def test_attention_mean_operand_equals_fp32_head_mean(dev):
    """The packed-P head mean leaving as a rollout operand == bf16(fp32 head mean of the same kernel), plus its row sum."""
    from vision_transformer_cam_b200 import ops
    for B, N, H in ((3, 197, 12), (2, 577, 16)):
        qkv = (_rand((B, N, 3 * H * 64), 90 + N, dev) * 1.5).bfloat16()
        _, _, mean = ops.attention_mean(qkv, H, 0.125)
        _, _, opnd = ops.attention_mean_operand(qkv, H, 0.125)
        ldr = ops.rollout_operand_ld(N)
        assert torch.equal(opnd[..., :N], mean.bfloat16())
        rs = opnd[..., ldr - 2:].contiguous().view(torch.float32)[..., 0]
        assert float((rs - mean.bfloat16().float().sum(-1)).abs().max()) < 1e-5 and float((rs - 1).abs().max()) < 5e-3


def test_cam_project_upsample_label(dev):
    from vision_transformer_cam_b200 import ops
    from oracle import postproc as PP
    B, N, D, Ccls = 4, 197, 768, 20
    X = _rand((B, N, D), 70, dev)
    W = _rand((Ccls, D), 71, dev, 0.1)
    cam = ops.cam_project(X, W)
    ref = PP.classic_cam(X.cpu(), W.cpu())
    assert float((cam.cpu() - ref).abs().max()) < 1e-5
    for hw in [(224, 224), (375, 500), (333, 401)]:
        up = ops.upsample_bilinear(cam, hw)
        ref_up = F.interpolate(ref, size=hw, mode="bilinear", align_corners=False)
        assert float((up.cpu() - ref_up).abs().max()) < 1e-5
        u8 = ops.upsample_bilinear(cam, hw, as_u8=True)
        ref_u8 = (ref_up * 255).to(torch.uint8)
        assert int((u8.cpu().int() - ref_u8.int()).abs().max()) <= 1
        labels = (torch.rand((B, Ccls), generator=torch.Generator().manual_seed(72)) < 0.15)
        labels[0] = False
        lab = ops.cam_label(cam, labels.to(dev), hw, 0.25)
        ref_lab = PP.cam_pseudo_label(ref, labels.float(), hw, 0.25)
        assert float((lab.cpu() == ref_lab).float().mean()) >= 0.9995      # argmax ties / 1e-6 interpolation noise
    m = cam.clone().view(B * Ccls, 196) + 0.1
    ops.normalize_max_(m)
    assert float((m.max(dim=1).values - 1).abs().max()) == 0.0


def test_hwp_vote_seg_confmat(dev):
    from vision_transformer_cam_b200 import ops
    from oracle import postproc as PP
    import numpy as np
    B, N, D, Ccls, K = 3, 197, 768, 20, 16
    X = _rand((B, N, D), 80, dev)
    W1 = _rand((Ccls, D), 81, dev, 0.3)
    idx = torch.stack([torch.randperm(196, generator=torch.Generator().manual_seed(82 + b))[:K] for b in range(B)])
    ori = torch.stack([X[b, 1 + idx[b].to(dev)] for b in range(B)]).contiguous()
    hwp = _rand((B, Ccls), 83, dev, 3.0)
    cls_rows = torch.rand((12, B, 12, N), generator=torch.Generator().manual_seed(84)).to(dev)
    cls_rows = cls_rows / cls_rows.sum(-1, keepdim=True)
    p2c, cos = ops.hwp_cos_vote(hwp, W1, ori, X)
    for b in range(B):
        ref_p2c = PP.hwp_patch_classes(hwp[b].cpu(), W1.cpu(), ori[b].cpu())
        mine = p2c[b].cpu().long()
        ok = (mine == ref_p2c) | ((mine < 0) & (ref_p2c >= 21))
        assert bool(ok.all()), (mine, ref_p2c)
        assert float((cos[b].cpu() - PP.hwp_cos_maps(X[b].cpu(), ori[b].cpu())).abs().max()) < 1e-5
    bgm = ops.cls_layer_map(cls_rows, 5, 12)
    hw = (375, 500)
    seg = ops.hwp_seg(cos, p2c, bgm, hw)
    ref_seg = PP.hwp_pseudo_seg(hwp.cpu(), W1.cpu(), ori.cpu(), X.cpu(), cls_rows.cpu(), hw, clamp_sentinel=True)
    assert float((seg.cpu() == ref_seg).float().mean()) >= 0.9995
    gt = torch.randint(0, 22, (B, *hw), generator=torch.Generator().manual_seed(85)).to(torch.uint8)
    gt[gt == 21] = 255
    mat = torch.zeros((21, 21), dtype=torch.int64, device=dev)
    ops.confmat_update(mat, gt.to(dev), seg)
    ref_mat = PP.confmat_update(None, gt.numpy(), seg.cpu().numpy())
    assert np.array_equal(mat.cpu().numpy(), ref_mat)            # integer counters: bit exact
    ops.confmat_update(mat, gt.to(dev), seg)
    assert np.array_equal(mat.cpu().numpy(), 2 * ref_mat)


# ------------------------------------------------------------------------- fp32 mode (split-bf16 operands)
def test_split_roundtrip_and_gemm_split(dev):
    """x ~= hi + lo carries 16 mantissa bits; the split GEMM (hi.hi + lo.hi + hi.lo, fp32 accumulation) must be ~2^-16
    accurate against an fp64 matmul of the SAME fp32 operands: 1e-5 of the output scale (bf16 GEMM: ~4e-3)."""
    from vision_transformer_cam_b200 import ops, _lib
    M, N, K = 1000, 768, 768
    a = _rand((M, K), 70, dev)
    w = _rand((N, K), 71, dev, 0.05)
    bias = _rand((N,), 72, dev)
    a2, w2 = ops.split_bf16(a), ops.split_bf16(w)
    assert a2.shape == (M, 2 * K)
    assert float((ops.merge_split(a2) - a).abs().max()) <= 2.0 ** -16 * float(a.abs().max())
    ref = (a.double() @ w.double().T + bias.double())
    out = ops.gemm_split(a2, w2, bias, _lib.EPI_BIAS)
    assert out.shape == (M, 2 * N)
    e = relerr(ops.merge_split(out), ref)
    assert e < 1e-5, e
    g = ops.gemm_split(a2, w2, bias, _lib.EPI_BIAS_GELU)
    eg = relerr(ops.merge_split(g), F.gelu(ref))
    assert eg < 1e-5, eg
    res = _rand((M, N), 73, dev)
    r = ops.gemm_split(a2, w2, bias, _lib.EPI_BIAS_RESIDUAL, residual=res)
    assert relerr(r, ref + res.double()) < 1e-5
    # bf16 GEMM on the same data for scale: three orders of magnitude worse
    assert relerr(ops.gemm_bf16(a.bfloat16(), w.bfloat16(), bias, _lib.EPI_BIAS).float(), ref) > 1e-4


def test_layernorm_and_patchify_split(dev):
    from vision_transformer_cam_b200 import ops
    x = _rand((300, 768), 74, dev, 3.0) + 0.5
    w, b = _rand((768,), 75, dev), _rand((768,), 76, dev)
    y = ops.layernorm_bf16(x, w, b, 1e-6, split=True)
    ref = F.layer_norm(x.double(), (768,), w.double(), b.double(), 1e-6)
    assert y.shape == (300, 1536) and relerr(ops.merge_split(y), ref) < 2e-5
    img = _rand((2, 3, 64, 64), 77, dev)
    p = ops.patchify(img, 16, split=True)
    refp = img.unfold(2, 16, 16).unfold(3, 16, 16).permute(0, 2, 3, 1, 4, 5).reshape(2 * 16, 768)
    assert p.shape == (32, 1536) and float((ops.merge_split(p) - refp).abs().max()) <= 2.0 ** -16 * float(refp.abs().max())


@pytest.mark.parametrize("B,N,H,masked", [(2, 197, 12, True), (1, 577, 4, True), (2, 130, 3, False)])
def test_attention_split(dev, B, N, H, masked):
    """fp32-mode attention: q, k, v and P as (hi, lo) pairs.  Against fp64 attention of the same fp32 q, k, v."""
    from vision_transformer_cam_b200 import ops
    D = H * 64
    qkv = _rand((B, N, 3 * D), 80, dev, 1.5)
    kb = None
    if masked:
        g = torch.Generator().manual_seed(81)
        kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
        kb[:, 0] = 0
        kb = kb.to(dev)
    q, k, v = qkv.double().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * 0.125
    if kb is not None:
        v_ = (kb != 0).double()
        s = s - 100.0 * torch.clamp(v_[:, :, None] + v_[:, None, :], max=1.0)[:, None]
    ref_p = s.softmax(-1)
    ref_o = (ref_p @ v).transpose(1, 2).reshape(B, N, D)
    out, cls, attn = ops.attention_kv(ops.split_bf16(qkv), H, 0.125, key_bias=kb, want_cls=True, want_attn=True, split=True)
    assert out.shape == (B, N, 2 * D)
    eo = relerr(ops.merge_split(out), ref_o)
    ep = float((attn.double() - ref_p).abs().max())
    print(f"split attention N={N}: O relerr {eo:.2e}, P abs err {ep:.2e}")
    assert eo < 3e-5, eo          # bf16 attention: ~5e-3
    assert ep < 2e-5 and float((cls.double() - ref_p[:, :, 0, :]).abs().max()) < 2e-5


# ----------------------------------------------------------------- the rows either side of the path
def test_patchify_u8_is_bit_identical_to_the_fp32_transform_path(dev):
    """uint8 HWC ingest: ToTensor (/255) + Normalize inside the im2col kernel == patchify(Normalize(ToTensor(img)))."""
    from vision_transformer_cam_b200 import ops
    g = torch.Generator().manual_seed(90)
    img = torch.randint(0, 256, (3, 64, 64, 3), generator=g, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    x = img.permute(0, 3, 1, 2).float().div(255)                       # ToTensor
    x = (x - torch.tensor(mean)[None, :, None, None]) / torch.tensor(std)[None, :, None, None]     # Normalize
    ref = ops.patchify(x.contiguous().to(dev), 16)
    out = ops.patchify_u8(img.to(dev), 16, mean, std)
    assert torch.equal(out, ref)
    out2 = ops.patchify_u8(img.to(dev), 16, mean, std, split=True)
    assert float((ops.merge_split(out2) - ops.merge_split(ops.patchify(x.contiguous().to(dev), 16, split=True))).abs().max()) == 0.0


def test_average_precision_matches_sklearn(dev):
    """On-device AP == sklearn.metrics.average_precision_score (what utils.py:258 calls), incl. tied scores and rows
    without positives (skipped by the reference, utils.py:256)."""
    from sklearn.metrics import average_precision_score
    from vision_transformer_cam_b200 import ops
    g = torch.Generator().manual_seed(91)
    B, C = 64, 20
    y = (torch.rand((B, C), generator=g) < 0.15).float()
    s = torch.rand((B, C), generator=g)
    s[:, ::3] = (s[:, ::3] * 4).round() / 4          # ties
    y[5] = 0
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    ap = ops.average_precision(y.to(dev), s.to(dev), acc).cpu()
    want = [average_precision_score(y[i].numpy(), s[i].numpy()) if y[i].sum() > 0 else -1.0 for i in range(B)]
    assert float((ap - torch.tensor(want, dtype=torch.float64)).abs().max()) < 1e-12
    valid = [w for w in want if w >= 0]
    assert abs(float(acc[0]) - sum(valid)) < 1e-9 and int(acc[1]) == len(valid)


def test_patch_similarity_reproduces_the_reference_normalize_quirk(dev):
    """predict.py:194-197: F.normalize(x) on [1,N,D] uses dim=1 (across tokens); sim = that . that^T."""
    from vision_transformer_cam_b200 import ops
    x = _rand((2, 197, 768), 92, dev)
    sim = ops.patch_similarity(x)
    for b in range(2):
        f = F.normalize(x[b:b + 1]).squeeze(0)
        ref = f.mm(f.t())
        assert float((sim[b] - ref).abs().max()) < 1e-5 * float(ref.abs().max())


# ------------------------------------------------------------------------------ LayerNorm fused into the GEMMs
@pytest.mark.parametrize("M", [300, 1000])
def test_layernorm_fused_gemms(dev, M):
    """x1 = x + a.Wp^T + bp (residual GEMM that also emits bf16(x1) and the row statistics), then
    gelu(LN(x1).W1^T + b1) evaluated on bf16(x1) with the normalisation folded into the epilogue -- against fp64 torch."""
    from vision_transformer_cam_b200 import ops
    D, HID = 768, 3072
    x = _rand((M, D), 100, dev, 1.5) + 0.3
    a = _rand((M, D), 101, dev).bfloat16()
    wp = _rand((D, D), 102, dev, 0.03).bfloat16()
    bp = _rand((D,), 103, dev, 0.1)
    x1_ref = x.double() + a.double() @ wp.double().T + bp.double()
    x1, x1b, stats = ops.gemm_resid_ln(a, wp, bp, x)
    assert relerr(x1, x1_ref) < 1e-5
    assert torch.equal(x1b, x1.bfloat16())
    s = stats.double().sum(1)
    assert relerr(s[:, 0], x1.double().sum(1)) < 1e-5 and relerr(s[:, 1], (x1.double() ** 2).sum(1)) < 1e-5
    # in place (out aliases residual) gives the same bits
    x_copy = x.clone()
    y1, y1b, st2 = ops.gemm_resid_ln(a, wp, bp, x_copy, out=x_copy)
    assert torch.equal(y1, x1) and torch.equal(y1b, x1b) and torch.equal(st2, stats)
    # residual_prep produces the same side outputs from an existing stream
    pb, pst = ops.residual_prep(x1)
    assert torch.equal(pb, x1b) and relerr(pst.double().sum(1), s) < 1e-6
    # folded LayerNorm + GEMM (+GELU)
    w1 = _rand((HID, D), 104, dev, 0.03)
    b1 = _rand((HID,), 105, dev, 0.1)
    gamma, beta = 1.0 + _rand((D,), 106, dev, 0.2), _rand((D,), 107, dev, 0.2)
    wf, g, c = ops.fold_ln(w1, gamma, beta, b1)
    assert torch.equal(wf, (w1 * gamma).bfloat16()) and relerr(g, wf.double().sum(1)) < 1e-5
    assert relerr(c, b1.double() + w1.double() @ beta.double()) < 1e-5
    ln = F.layer_norm(x1.double(), (D,), gamma.double(), beta.double(), 1e-6)
    for gelu in (False, True):
        ref = ln @ w1.double().T + b1.double()
        ref = F.gelu(ref) if gelu else ref
        out = ops.gemm_lnfold(x1b, wf, c, g, stats, 1e-6, gelu=gelu)
        # the unfused path for scale: LayerNorm kernel -> bf16 GEMM
        base = ops.gemm_bf16(ops.layernorm_bf16(x1, gamma, beta, 1e-6), w1.bfloat16(), b1, 1 if gelu else 0)
        e, e0 = relerr(out.float(), ref), relerr(base.float(), ref)
        print(f"LN-folded GEMM gelu={gelu}: relerr {e:.2e} (separate LayerNorm kernel: {e0:.2e})")
        assert e < 8e-3 and e < 2.0 * e0 + 1e-3


@pytest.mark.parametrize("B,N,H,masked", [(3, 197, 12, True), (2, 197, 12, False), (5, 50, 3, True), (2, 208, 2, False), (300, 197, 1, False),
                                          (2, 577, 16, True), (1, 785, 12, False), (2, 300, 3, True), (1, 1030, 2, True)])
def test_attention_fused_head_mean(dev, B, N, H, masked):
    """Head mean of P from the packed bf16 P of the fast attention kernel (vtc_attention_mean: the rollout's input,
    predict.py:189-190) == mean over heads of the fp32 softmax; P enters as the bf16 values the P.V product uses, so the
    bar is bf16-level: 1 % of the largest entry."""
    from vision_transformer_cam_b200 import ops
    qkv = _rand((B, N, 3 * H * 64), 110, dev, 1.5).bfloat16()
    kb = None
    if masked:
        g = torch.Generator().manual_seed(111)
        kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
        kb[:, 0] = 0
        kb = kb.to(dev)
    ref_o, ref_p = _attn_ref(qkv, H, 0.125, kb)
    out, cls, mean = ops.attention_mean(qkv, H, 0.125, key_bias=kb)
    ref_m = ref_p.mean(1)
    assert float((mean - ref_m).abs().max()) <= 1e-2 * float(ref_m.max()), float((mean - ref_m).abs().max())
    assert float((mean.sum(-1) - 1).abs().max()) < 5e-3
    assert relerr(out.float(), ref_o) < 1e-2 and float((cls - ref_p[:, :, 0, :]).abs().max()) < 5e-6
    out2, cls2, mean2 = ops.attention_mean(qkv, H, 0.125, key_bias=kb)
    assert torch.equal(mean, mean2) and torch.equal(out, out2)            # fixed accumulation order


@pytest.mark.parametrize("N,amp", [(197, 12.0), (577, 6.0), (785, 12.0)])
def test_attention_fused_head_mean_large_dynamic_range(dev, N, amp):
    """Peaked logits: the running maximum rises by many octaves inside a row (within a key block and from block to block),
    so chunks of the packed P are stored against different reference maxima; the head mean must still be the softmax's."""
    from vision_transformer_cam_b200 import ops
    B, H = 2, 4
    qkv = _rand((B, N, 3 * H * 64), 120, dev, 1.0)
    qkv[:, :, : 2 * H * 64] *= amp
    qkv = qkv.bfloat16()
    g = torch.Generator().manual_seed(121)
    kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
    kb[:, 0] = 0
    kb = kb.to(dev)
    for bias in (None, kb):
        ref_o, ref_p = _attn_ref(qkv, H, 0.125, bias)
        out, cls, mean = ops.attention_mean(qkv, H, 0.125, key_bias=bias)
        ref_m = ref_p.mean(1)
        assert float((mean - ref_m).abs().max()) <= 1e-2 * float(ref_m.max()), float((mean - ref_m).abs().max())
        assert float((mean.sum(-1) - 1).abs().max()) < 5e-3
        assert relerr(out.float(), ref_o) < 1.5e-2


def _attn_ref_hd(qkv, H, scale, kb):
    """_attn_ref for any head dimension."""
    B, N, D3 = qkv.shape
    D = D3 // 3
    q, k, v = qkv.float().view(B, N, 3, H, D // H).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * scale
    if kb is not None:
        v_ = (kb != 0).float()
        s = s - 100.0 * torch.clamp(v_[:, :, None] + v_[:, None, :], max=1.0)[:, None]
    p = s.softmax(-1)
    return (p @ v).transpose(1, 2).reshape(B, N, D), p


@pytest.mark.parametrize("B,N,H,hd,masked", [(2, 257, 3, 80, True), (1, 257, 16, 80, False), (3, 100, 2, 96, True), (2, 33, 2, 16, False),
                                             (1, 320, 1, 128, True), (2, 197, 2, 64, True)])
def test_attention_generic_head_dims(dev, B, N, H, hd, masked):
    """The general-shape attention path (ViT-H/14: head_dim 80, 257 tokens) against the fp32 softmax attention of the same
    bf16 operands: P (fp32, not rounded before P V here) to 2e-6, O to the bf16 rounding of the output."""
    from vision_transformer_cam_b200 import ops
    qkv = _rand((B, N, 3 * H * hd), 130, dev, 1.2).bfloat16()
    kb = None
    if masked:
        g = torch.Generator().manual_seed(131)
        kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0)
        kb[:, 0] = 0
        kb = kb.to(dev)
    scale = hd ** -0.5
    out, cls, attn = ops.attention_generic(qkv, H, scale, key_bias=kb, want_cls=True, want_attn=True)
    ref_o, ref_p = _attn_ref_hd(qkv, H, scale, kb)
    assert float((attn - ref_p).abs().max()) < 5e-6, float((attn - ref_p).abs().max())
    assert torch.equal(cls, attn[:, :, 0, :])
    assert relerr(out.float(), ref_o) < 5e-3, relerr(out.float(), ref_o)
    out2, cls2, none = ops.attention_generic(qkv, H, scale, key_bias=kb, want_cls=True, want_attn=False)
    assert none is None and torch.equal(out, out2) and torch.equal(cls, cls2)
    if hd == 64:      # same operator as the tensor-core kernels
        o64, c64, _ = ops.attention(qkv, H, scale, key_bias=kb)
        assert relerr(o64.float(), out.float()) < 1e-2 and float((c64 - cls).abs().max()) < 5e-6
