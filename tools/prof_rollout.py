"""Per-kernel time of the forward: plain, with every layer's fp32 head-mean P (attn_mean), and with the fused rollout output
(bf16 rollout operands + the streaming rollout kernel).    python tools/prof_rollout.py [batch]"""
import sys, torch
sys.path.insert(0, ".")
import vision_transformer_cam_b200 as V
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
x = torch.randn((B, 3, 224, 224), device=dev)
for name, kw in (("plain", {}), ("attn_mean", dict(attn_mean=True)), ("rollout", dict(rollout=True))):
    for _ in range(2):
        model.forward_cam(x, **kw)
    model.kernel_profile(True)
    for _ in range(3):
        model.forward_cam(x, **kw)
    prof = model.kernel_profile()
    model.kernel_profile(False)
    print(name, {k: round(v[0] / 3, 3) for k, v in prof.items() if v[1]}, "total", round(sum(v[0] for v in prof.values()) / 3, 2), flush=True)
