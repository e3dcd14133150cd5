"""BASELINE config 3: CAM (+ optionally attention rollout) extraction over a synthetic train_aug-sized set (10,582 images),
batch-sharded over the GPUs of one box with an NCCL gather of the maps (pipeline.extract_cams_sharded).

    python tools/run_config3.py [--n 10582] [--batch 256] [--rollout]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config3.py

Images are generated on the device per batch (synthetic N(0,1), seed = global image index block); prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import dist as D, pipeline

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10582)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--rollout", action="store_true")
args = ap.parse_args()
rank, local_rank, world = D.init_from_env()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()


def get(lo, hi):
    g = torch.Generator(device=dev).manual_seed(1000 + lo)
    return torch.randn((hi - lo, 3, 224, 224), generator=g, device=dev)


pipeline.extract_cams_sharded(model, get, n_items=min(args.n, 2 * args.batch * world), batch=args.batch, with_rollout=args.rollout)   # warm-up
if world > 1:
    torch.distributed.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
out = pipeline.extract_cams_sharded(model, get, n_items=args.n, batch=args.batch, with_rollout=args.rollout)
torch.cuda.synchronize()
dt = D.max_over_ranks(time.perf_counter() - t0, dev)
if rank == 0:
    print(json.dumps({"config": "ViT-B/16 224px CAM%s over %d synthetic images, batch-sharded over %d GPU(s), NCCL gather" %
                      (" + rollout" if args.rollout else "", args.n, world), "images": args.n, "n_gpus": world, "seconds": dt,
                      "images_per_s": args.n / dt, "cam_shape": list(out["cam"].shape), "finite": bool(torch.isfinite(out["cam"]).all())}), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
