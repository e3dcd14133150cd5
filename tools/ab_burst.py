"""A/B helper: median / min step time of the B=256 forward + CAM over short bursts separated by idle gaps (so that the
power cap does not blur small differences).  Compare runs with different environment toggles (VTC_NO_PDL, ...)."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import cam as CAM
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
x = torch.randn((B, 3, 224, 224), device=dev)
graphed = os.environ.get("AB_GRAPH") == "1"
def step():
    o = model.forward_cam_graphed(x) if graphed else model.forward_cam(x)
    return CAM.classic_cam(o.tokens_last, model.head1.weight.data)
for _ in range(5):
    step()
torch.cuda.synchronize()
res = []
for burst in range(8):
    time.sleep(0.7)
    step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        step()
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 4)
print(f"{os.environ.get('AB_TAG', '')}: median {statistics.median(res):.3f} ms  min {min(res):.3f} ms  ({B / statistics.median(res) * 1e3:.0f} images/s)")
