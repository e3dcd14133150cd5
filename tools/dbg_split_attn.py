import sys, torch
sys.path.insert(0, ".")
from vision_transformer_cam_b200 import ops
dev = "cuda:0"
def run(B, N, H, masked, scale_in=1.5):
    D = H * 64
    g = torch.Generator().manual_seed(80)
    qkv = (torch.randn((B, N, 3 * D), generator=g) * scale_in).to(dev)
    kb = None
    if masked:
        g = torch.Generator().manual_seed(81)
        kb = torch.where(torch.rand((B, N), generator=g) < 0.3, -100.0, 0.0); kb[:, 0] = 0; kb = kb.to(dev)
    q, k, v = qkv.double().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-2, -1)) * 0.125
    if kb is not None:
        v_ = (kb != 0).double()
        s = s - 100.0 * torch.clamp(v_[:, :, None] + v_[:, None, :], max=1.0)[:, None]
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, D)
    out, cls, attn = ops.attention_kv(ops.split_bf16(qkv), H, 0.125, key_bias=kb, want_cls=True, want_attn=True, split=True)
    om = ops.merge_split(out).double()
    err = (om - o).abs()
    rows = err.amax(dim=(0, 2))
    print(f"N={N} H={H} masked={masked}: O relerr {float(err.max()/o.abs().max()):.2e} P err {float((attn.double()-p).abs().max()):.2e} worst rows {rows.topk(5).indices.tolist()} mean err {float(err.mean()):.2e}")
    if kb is not None:
        bgrow = (kb != 0)
        e_bg = err[bgrow].max() if bgrow.any() else 0
        e_fg = err[~bgrow].max()
        print(f"   err on background query rows {float(e_bg):.2e}, foreground rows {float(e_fg):.2e}")
for N in (130, 197, 256):
    for m in (False, True):
        run(2, N, 4, m)
run(2, 197, 4, True, 0.5)
