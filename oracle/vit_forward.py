"""Oracle (test infrastructure): fp32 CPU restatement of the reference ViT forward.

Follows /root/reference/vit_model.py op-for-op so that, on ViT-B/16-224 where the
reference runs, the outputs are bit-identical on the same torch build (checked by
`tests/golden/make_golden.py`, result recorded in `tests/golden/README.md`).

Only two things are generalised: the hard-coded `197`
(`vit_model.py:319,350`) becomes N and the hard-coded `12` heads (`:123`) becomes H.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class VitConfig:
    img_size: int = 224
    patch_size: int = 16
    in_c: int = 3
    num_classes: int = 20
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: float = 4.0
    representation_size: Optional[int] = None
    mask_from: int = 4          # vit_model.py:118,325 (absolute layer index)
    mask_thresh: float = 0.25   # vit_model.py:339
    topk: int = 16              # vit_model.py:377
    ln_eps: float = 1e-6        # vit_model.py:244

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def num_tokens(self) -> int:
        return self.num_patches + 1

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)


VIT_B16_224 = VitConfig()
VIT_B16_448 = VitConfig(img_size=448)
VIT_L16_384 = VitConfig(img_size=384, embed_dim=1024, depth=24, num_heads=16)


def init_state_dict(cfg: VitConfig = VIT_B16_224, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Reproduce the reference constructor's RNG stream (vit_model.py:215-301, 427-442).

    Construction order (default nn init consumes RNG), then trunc_normal(pos_embed, cls_token),
    then `apply(_init_vit_weights)` in child-first registration order, then `norm1`, `norm2`,
    `head1` created afterwards with the default nn.Linear / LayerNorm init (`:292-295`).
    """
    torch.manual_seed(seed)
    D, C = cfg.embed_dim, cfg.num_classes
    sd: Dict[str, torch.Tensor] = {}
    conv = nn.Conv2d(cfg.in_c, D, kernel_size=cfg.patch_size, stride=cfg.patch_size)   # :64
    cls_token = torch.zeros(1, 1, D)                                                     # :249
    pos_embed = torch.zeros(1, cfg.num_tokens, D)                                        # :251
    blocks = []
    for _ in range(cfg.depth):                                                           # :255-260
        blk = dict(norm1=nn.LayerNorm(D, eps=cfg.ln_eps), qkv=nn.Linear(D, 3 * D, bias=True),
                   proj=nn.Linear(D, D), norm2=nn.LayerNorm(D, eps=cfg.ln_eps),
                   fc1=nn.Linear(D, cfg.hidden), fc2=nn.Linear(cfg.hidden, D))
        blocks.append(blk)
    norm = nn.LayerNorm(D, eps=cfg.ln_eps)                                               # :262
    pre_fc = None
    feat = D
    if cfg.representation_size:                                                          # :267-273
        pre_fc = nn.Linear(D, cfg.representation_size)
        feat = cfg.representation_size
    head = nn.Linear(feat, C)                                                            # :279
    nn.init.trunc_normal_(pos_embed, std=0.02)                                           # :285
    nn.init.trunc_normal_(cls_token, std=0.02)                                           # :289

    def _init(m):                                                                        # :427-442
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.01)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out")
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)

    _init(conv)
    for blk in blocks:
        for name in ("norm1", "qkv", "proj", "norm2", "fc1", "fc2"):
            _init(blk[name])
    _init(norm)
    if pre_fc is not None:
        _init(pre_fc)
    _init(head)
    norm1 = nn.LayerNorm(256, eps=cfg.ln_eps)                                            # :292
    norm2 = nn.LayerNorm(32, eps=cfg.ln_eps)                                             # :293
    head1 = nn.Linear(feat, C)                                                           # :295

    sd["cls_token"] = cls_token
    sd["pos_embed"] = pos_embed
    sd["patch_embed.proj.weight"] = conv.weight.detach()
    sd["patch_embed.proj.bias"] = conv.bias.detach()
    for i, blk in enumerate(blocks):
        p = f"blocks.{i}."
        for mod, key in (("norm1", "norm1"), ("qkv", "attn.qkv"), ("proj", "attn.proj"),
                         ("norm2", "norm2"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2")):
            sd[p + key + ".weight"] = blk[mod].weight.detach()
            sd[p + key + ".bias"] = blk[mod].bias.detach()
    sd["norm.weight"], sd["norm.bias"] = norm.weight.detach(), norm.bias.detach()
    if pre_fc is not None:
        sd["pre_logits.fc.weight"], sd["pre_logits.fc.bias"] = pre_fc.weight.detach(), pre_fc.bias.detach()
    sd["head.weight"], sd["head.bias"] = head.weight.detach(), head.bias.detach()
    sd["norm1.weight"], sd["norm1.bias"] = norm1.weight.detach(), norm1.bias.detach()
    sd["norm2.weight"], sd["norm2.bias"] = norm2.weight.detach(), norm2.bias.detach()
    sd["head1.weight"], sd["head1.bias"] = head1.weight.detach(), head1.bias.detach()
    return {k: v.clone() for k, v in sd.items()}


def peaked(sd: Dict[str, torch.Tensor], qkv_scale: float = 5.0, head1_scale: float = 30.0,
           head1_bias: float = 1.0) -> Dict[str, torch.Tensor]:
    """The 'peaked' parity regime of SURVEY.md section 8(d): qkv weights x5 so that the layer>=4
    background mask fires (fraction 0.24-0.40), head1 x30 (+1 bias) so sigmoid>=0.9 fires."""
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        if k.endswith("attn.qkv.weight"):
            out[k] *= qkv_scale
    out["head1.weight"] *= head1_scale
    out["head1.bias"] += head1_bias
    return out


def masked(sd: Dict[str, torch.Tensor], qk_scale: float = 3.5, first_layer: int = 3, head1_scale: float = 30.0,
           head1_bias: float = 1.0) -> Dict[str, torch.Tensor]:
    """The 'masked' parity regime: the q and k rows of `attn.qkv.weight` (v untouched) of blocks >= `first_layer`
    are scaled by 3.5, head1 as in `peaked`.  The attention logits stay at a natural scale (|S| of a few units, the
    residual stream is not swamped by the attention output as in `peaked`), yet the CLS row varies enough for the
    layer>4 background mask of vit_model.py:325-361 to fire on 0.5-0.6 of the patches (layers 5..11; asserted to lie
    in (0.1, 0.9) where it is used).  A CPU emulation of bf16-operand arithmetic (tools/emulate_bf16.py) sits at
    6.5e-3 of the fp32 logits here, i.e. the north-star 1e-2 bar is meaningful in this regime."""
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        if k.endswith("attn.qkv.weight") and int(k.split(".")[1]) >= first_layer:
            two_d = 2 * out[k].shape[1]
            out[k][:two_d] *= qk_scale
    out["head1.weight"] *= head1_scale
    out["head1.bias"] += head1_bias
    return out


# the batch-global max of vit_model.py:335 grows with the batch (one outlier patch among 256 x 196), so a background
# fraction inside (0.1, 0.9) (0.36-0.40 on layers 5..11) needs a slightly smaller q/k scale at B = 256 than at B = 3
MASKED_B256_QK_SCALE = 3.0


def make_images(first_index: int, count: int, size: int = 224, device="cpu") -> torch.Tensor:
    """Synthetic ImageNet-normalised images: N(0,1), seed 1000+global index."""
    imgs = []
    for i in range(first_index, first_index + count):
        g = torch.Generator().manual_seed(1000 + i)
        imgs.append(torch.randn(3, size, size, generator=g))
    return torch.stack(imgs).to(device)


def cls_row_stat(P: torch.Tensor) -> torch.Tensor:
    """vit_model.py:329-335 / 366-372 without the final `/ max`: head-mean attention, + identity,
    renormalise by the row sum, CLS row, patch columns.  Returns [B, N-1]."""
    att = torch.mean(P, dim=1)                                          # :329
    eye = torch.eye(att.size(2), dtype=att.dtype, device=att.device)   # :331
    aug = att + eye                                                     # :332
    aug = aug / aug.sum(dim=-1).unsqueeze(-1)                           # :333
    return aug[:, 0, 1:]                                                # :334


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: VitConfig = VIT_B16_224,
            mask_norm: str = "batch", keep_P: bool = True,
            forced_bg: Optional[List[Optional[torch.Tensor]]] = None,
            forced_topk: Optional[torch.Tensor] = None) -> Dict[str, object]:
    """Reference forward (vit_model.py:303-424), fp32, any device.

    Returns a dict: logits [B,C]; P list_L [B,H,N,N] (if keep_P); X list_L [B,N,D]; hwp [B,C];
    ori [B,16,D]; topk_idx [B,16]; cls_rows [L,B,H,N] (P[:, :, 0, :]); bg list (per layer the
    {0,1} background vector fed to layer l+1, or None); c_last [B,N-1].

    `forced_bg` / `forced_topk` teacher-force the discrete decisions (SURVEY.md hard part 4) so
    continuous outputs can be compared in low precision without decision flips cascading.
    """
    assert mask_norm in ("batch", "image")
    B = x.shape[0]
    D, H, hd, N, L = cfg.embed_dim, cfg.num_heads, cfg.head_dim, cfg.num_tokens, cfg.depth
    assert x.shape[2] == cfg.img_size and x.shape[3] == cfg.img_size                   # :69-70
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch_size)  # :76
    t = t.flatten(2).transpose(1, 2)                                                    # :79
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)                        # :308-310
    t = t + sd["pos_embed"]                                                             # :314
    scale = hd ** -0.5                                                                  # :97
    mask = None                      # additive [B,N,N] bias for the NEXT layer
    P_list, X_list, bg_list, cls_rows = [], [], [], []
    P = None
    for l in range(L):                                                                  # :320
        p = f"blocks.{l}."
        y = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps)          # :193
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])                         # :110
        qkv = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]                                                              # :113
        attn = (q @ k.transpose(-2, -1)) * scale                                                      # :119/:122
        if l > cfg.mask_from and mask is not None:                                                    # :118-124
            attn = attn + mask.unsqueeze(1)
        attn = attn.softmax(dim=-1)                                                                   # :126
        P = attn
        o = (attn @ v).transpose(1, 2).reshape(B, N, D)                                               # :135-137
        o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])                         # :138
        t = t + o                                                                                     # :194
        y2 = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.ln_eps)
        h = F.gelu(F.linear(y2, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))                    # :158-159
        t = t + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])                         # :161,:198
        if keep_P:
            P_list.append(P)
        cls_rows.append(P[:, :, 0, :].clone())
        X_list.append(t)                                                                              # :324
        bg = None
        if l >= cfg.mask_from:                                                                        # :325
            m = cls_row_stat(P)                                                                       # :329-334
            if mask_norm == "batch":
                m14 = m / m.max()                                                                     # :335 (GLOBAL max)
            else:
                m14 = m / m.max(dim=1, keepdim=True).values
            bg = torch.lt(m14, cfg.mask_thresh).to(t.dtype)                                           # :337-342
            if forced_bg is not None and forced_bg[l] is not None:
                bg = forced_bg[l].to(t.dtype)
            vfull = torch.cat((torch.zeros(B, 1, dtype=t.dtype, device=t.device), bg), dim=1)         # :347-348
            mm = vfull.unsqueeze(2).repeat(1, 1, N)                                                   # :350
            mm = mm + mm.permute(0, 2, 1)                                                             # :351
            mm[mm > 1] = 1                                                                            # :353
            mask = -100 * mm                                                                          # :361
        bg_list.append(bg)
    m = cls_row_stat(P)                                                                               # :366-371
    m14 = m / m.max() if mask_norm == "batch" else m / m.max(dim=1, keepdim=True).values              # :372
    if forced_topk is not None:
        idx = forced_topk.to(torch.long)
    else:
        idx = torch.stack([torch.topk(m14[j], cfg.topk, dim=0).indices for j in range(B)])            # :377
    ori = torch.stack([t[j][idx[j] + 1] for j in range(B)])                                           # :381-390
    hwp = F.linear(ori.mean(dim=1), sd["head1.weight"], sd["head1.bias"])                             # :392-393
    xn = F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps)                        # :402
    feat = xn[:, 0]
    if "pre_logits.fc.weight" in sd:                                                                  # :269-273
        feat = torch.tanh(F.linear(feat, sd["pre_logits.fc.weight"], sd["pre_logits.fc.bias"]))
    logits = F.linear(feat, sd["head.weight"], sd["head.bias"])                                       # :422
    return dict(logits=logits, P=P_list, X=X_list, hwp=hwp, clsh1_weight=sd["head1.weight"], ori=ori,
                topk_idx=idx, cls_rows=torch.stack(cls_rows), bg=bg_list, c_last=m)


def as_reference_tuple(out: Dict[str, object], cfg: VitConfig = VIT_B16_224):
    """The 6-tuple of vit_model.py:424; for depth>12 only the last 12 layers are kept (:322)."""
    keep = slice(max(0, cfg.depth - 12), None)
    return (out["logits"], out["P"][keep], out["X"][keep], out["hwp"], out["clsh1_weight"], out["ori"])
