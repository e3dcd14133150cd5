"""Per-kernel time of the forward when every layer's head-mean P is requested (rollout input)."""
import sys, torch
sys.path.insert(0, ".")
import vision_transformer_cam_b200 as V
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
x = torch.randn((B, 3, 224, 224), device=dev)
for am in (False, True):
    for _ in range(2):
        model.forward_cam(x, attn_mean=am)
    model.kernel_profile(True)
    for _ in range(3):
        model.forward_cam(x, attn_mean=am)
    prof = model.kernel_profile()
    model.kernel_profile(False)
    print("attn_mean", am, {k: round(v[0] / 3, 3) for k, v in prof.items() if v[1]}, "total", round(sum(v[0] for v in prof.values()) / 3, 2))
