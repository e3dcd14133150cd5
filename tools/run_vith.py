"""ViT-H/14-224 (general-shape path: padded patch GEMM, FMA-pipe attention) forward + CAM on one GPU: ms / step, images/s."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import cam as CAM
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = V.vit_huge_patch14_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
x = torch.randn((B, 3, 224, 224), device=dev)
def step():
    o = model.forward_cam(x)
    return CAM.classic_cam(o.tokens_last, model.head1.weight.data)
for _ in range(2):
    step()
model.kernel_profile(True)
for _ in range(3):
    step()
prof = model.kernel_profile()
model.kernel_profile(False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"case": "ViT-H/14-224 forward + CAM", "batch": B, "ms_per_step": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1),
                  "kernels_ms": {k: round(v[0] / 3, 2) for k, v in prof.items() if v[1]}}))
