"""BASELINE configs 4 and 5 on one GPU (per-GPU figures; config 5's 8-GPU data parallelism replicates this per rank):
  4: ViT-B/16 448 px (785 tokens), B = 64, forward + CAM + attention rollout through all 12 layers
  5: ViT-L/16 384 px (577 tokens, 24 layers), B = 64, forward + CAM
Synthetic N(0,1) images generated on the device, random-init weights; CUDA-event timing, prints one JSON line per case.

    python tools/run_configs45.py [--steps 10] [--batch 64]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import cam as CAM

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=64)
args = ap.parse_args()
dev = torch.device("cuda:0")


def gflop_per_image(S, patch, D, L, H, C):
    P = (S // patch) ** 2
    N = P + 1
    hd = D // H
    return (2 * P * 3 * patch * patch * D + L * (2 * N * D * 3 * D + 4 * H * N * N * hd + 2 * N * D * D + 16 * N * D * D) + 4 * D * C + 2 * P * D * C) / 1e9


cases = [("config 4: ViT-B/16-448 forward + CAM + rollout", dict(img_size=448, patch_size=16, embed_dim=768, depth=12, num_heads=12), True),
         ("config 4 without the rollout", dict(img_size=448, patch_size=16, embed_dim=768, depth=12, num_heads=12), False),
         ("config 5: ViT-L/16-384 forward + CAM", dict(img_size=384, patch_size=16, embed_dim=1024, depth=24, num_heads=16), False)]
for name, kw, rollout in cases:
    torch.manual_seed(0)
    model = V.VisionTransformer(num_classes=20, representation_size=None, **kw).to(dev).eval()
    x = torch.randn((args.batch, 3, kw["img_size"], kw["img_size"]), device=dev)

    def step():
        o = model.forward_cam(x, rollout=rollout)
        cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
        return (cam, o.rollout) if rollout else (cam,)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    gf = gflop_per_image(kw["img_size"], 16, kw["embed_dim"], kw["depth"], kw["num_heads"], 20)
    print(json.dumps({"case": name, "batch": args.batch, "ms_per_step": round(ms, 3), "images_per_s": round(args.batch / ms * 1e3, 1),
                      "gflop_per_image": round(gf, 3), "tflops": round(args.batch * gf / ms, 1), "finite": bool(all(torch.isfinite(t).all() for t in out))}), flush=True)
    del model, x
    torch.cuda.empty_cache()
