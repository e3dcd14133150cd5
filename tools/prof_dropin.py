"""Latency / throughput of the reference-facing calls (not the bench workload):
  * the reference-compatible 6-tuple `model(x)` (full P [B,H,N,N] and tokens of 12 layers materialised) at B = 1, 32, 256
  * `pipeline.predict` (predict.py:129-293 for a batch) at B = 1
  * `forward_cam` at B = 1 (launch-bound)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import pipeline
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for B in (1, 32, 256):
    x = torch.randn((B, 3, 224, 224), device=dev)
    ms = timed(lambda: model(x), 20 if B < 256 else 5)
    print(f"6-tuple model(x)      B={B:3d}: {ms:8.3f} ms  {B / ms * 1e3:9.1f} images/s")
    ms = timed(lambda: model.forward_cam(x), 50 if B < 256 else 10)
    print(f"forward_cam           B={B:3d}: {ms:8.3f} ms  {B / ms * 1e3:9.1f} images/s")
    if hasattr(model, "forward_cam_graphed"):
        ms = timed(lambda: model.forward_cam_graphed(x), 50 if B < 256 else 10)
        print(f"forward_cam (graph)   B={B:3d}: {ms:8.3f} ms  {B / ms * 1e3:9.1f} images/s")
x = torch.randn((1, 3, 224, 224), device=dev)
ms = timed(lambda: pipeline.predict(model, x, (375, 500)), 50)
print(f"pipeline.predict      B=  1: {ms:8.3f} ms (rollout + 12 layer maps at 375x500 + CAM)")
