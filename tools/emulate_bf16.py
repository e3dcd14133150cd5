"""CPU emulation of the bf16 rounding points of the CUDA path (operands bf16, fp32 accumulate), to budget the
expected deviation from the fp32 oracle.  Test infrastructure (imports oracle/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import vit_forward as VF


def r(x, on=True):
    return x.bfloat16().float() if on else x


def forward_emul(sd, x, cfg, forced_bg, forced_topk, flags):
    B = x.shape[0]
    D, H, hd, N, L = cfg.embed_dim, cfg.num_heads, cfg.head_dim, cfg.num_tokens, cfg.depth
    f = lambda name: flags.get(name, True)
    W = lambda k: r(sd[k], f("w"))
    t = F.conv2d(r(x, f("x")), W("patch_embed.proj.weight"), sd["patch_embed.proj.bias"], stride=cfg.patch_size)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1) + sd["pos_embed"]
    kb = None
    for l in range(L):
        p = f"blocks.{l}."
        y = r(F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps), f("ln"))
        qkv = r(F.linear(y, W(p + "attn.qkv.weight"), sd[p + "attn.qkv.bias"]), f("qkv"))
        qkv = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        s = (q @ k.transpose(-2, -1)) * hd ** -0.5
        if l > cfg.mask_from and kb is not None:
            v_ = (kb != 0).float()
            s = s - 100.0 * torch.clamp(v_[:, :, None] + v_[:, None, :], max=1.0)[:, None]
        m = s.max(-1, keepdim=True).values
        e = torch.exp(s - m)
        denom = e.sum(-1, keepdim=True)
        o = (r(e, f("p")) @ v) / denom
        o = r(o.transpose(1, 2).reshape(B, N, D), f("ao"))
        t = t + F.linear(o, W(p + "attn.proj.weight"), sd[p + "attn.proj.bias"])
        y2 = r(F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.ln_eps), f("ln"))
        h = r(F.gelu(F.linear(y2, W(p + "mlp.fc1.weight"), sd[p + "mlp.fc1.bias"])), f("h"))
        t = t + F.linear(h, W(p + "mlp.fc2.weight"), sd[p + "mlp.fc2.bias"])
        if l >= cfg.mask_from:
            bg = forced_bg[l]
            kb = torch.cat((torch.zeros(B, 1), bg), dim=1) * -100.0
    xn = F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps)
    return F.linear(xn[:, 0], sd["head.weight"], sd["head.bias"]), t


if __name__ == "__main__":
    cfg = VF.VIT_B16_224
    sd0 = VF.init_state_dict(cfg, 0)
    for regime, sd in (("default", sd0), ("peaked", VF.peaked(sd0))):
        x = VF.make_images(0, 3)
        ref = VF.forward(sd, x, cfg, keep_P=False)
        fb = ref["bg"]
        for name, flags in (("all", {}), ("no-qkv", {"qkv": False}), ("no-p", {"p": False}), ("no-w", {"w": False}),
                            ("no-ln", {"ln": False}), ("no-h", {"h": False}), ("no-ao", {"ao": False}), ("none", {k: False for k in ("w", "x", "ln", "qkv", "p", "ao", "h")})):
            lg, t = forward_emul(sd, x, cfg, fb, None, flags)
            e = float((lg - ref["logits"]).abs().max() / ref["logits"].abs().max())
            et = float((t - ref["X"][-1]).abs().max() / ref["X"][-1].abs().max())
            print(f"{regime:8s} {name:7s} logits relerr {e:.2e}  tokens relerr {et:.2e}")
