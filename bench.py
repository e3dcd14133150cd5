#!/usr/bin/env python
"""Headline benchmark: images/s of the ViT-B/16 224 px forward + patch-token CAM (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference] [--no-extras]

One "step" = one pass of the hot path over one batch of B synthetic images per GPU: fused forward (patch embed, 12
blocks with the background mask, hwp head, final norm + head) + classic CAM projection (20 x 14 x 14 maps per image).
Prints ONE JSON line (rank 0).  `value` is whole-job images/s with the inputs resident in HBM; `e2e` the same metric with
pinned-host inputs copied H2D and the CAMs + logits copied D2H inside the timed region; `roofline` the dominant kernel
(the tcgen05 GEMM) timed live with CUDA events; `cpu_baseline` the oracle port of the reference on the host cores.
Under torchrun (N > 1) every rank processes its own batch (weak scaling); the per-step CAMs are all-gathered and the
counters all-reduced over NCCL inside the timed region; time = max over ranks.

Beside the headline the line carries (all measured in this run):
  `parity`              logits relerr / CAM cosine / pseudo-label agreement of the GPU arm against the CPU-baseline outputs on
                        the same 256 images and weights (the "CAM agreement" third of BASELINE.json:metric), N = 1 only;
  `gpu_eager_baseline`  the reference model itself (oracle/_ref, else the oracle port) run eagerly on the same B200 at B = 256
                        under torch.no_grad(): fp32 with TF32 off, and torch.autocast(bfloat16), N = 1 only;
  `configs`             BASELINE configs 3 (10,582 images, CAM + rollout, strong scaling over the ranks, NCCL gather), 4
                        (ViT-B/16 448 px, rollout through 12 layers) and 5 (ViT-L/16 384 px), a few steps each.
The reference arm (`--impl reference`) times the reference's own CPU implementation (oracle/_ref/vit_model.py, unmodified,
through oracle/ref_shim.py; the oracle port only if that copy is missing) at the batch it prints.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG, PATCH, D, L, H, HID, C = 224, 16, 768, 12, 12, 3072, 20
P = (IMG // PATCH) ** 2
N = P + 1
# algorithmic FLOPs per image: 2*MAC, no padding, no recompute
FLOP_PATCH = 2 * P * (3 * PATCH * PATCH) * D
FLOP_QKV = 2 * N * D * 3 * D
FLOP_ATTN = 4 * H * N * N * (D // H)
FLOP_PROJ = 2 * N * D * D
FLOP_FC = 2 * N * D * HID
FLOP_IMAGE = FLOP_PATCH + L * (FLOP_QKV + FLOP_ATTN + FLOP_PROJ + 2 * FLOP_FC) + 2 * (2 * D * C) + 2 * P * D * C
GEMM_FLOP_IMAGE = FLOP_PATCH + L * (FLOP_QKV + FLOP_PROJ + 2 * FLOP_FC)


TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic.json")


def gemm_traffic():
    """DRAM bytes per GEMM launch (launch-weighted mean over the qkv / proj / fc1 / fc2 GEMMs of a layer) from the committed
    ncu --set full capture (profiles/r02_traffic.json, else r01); None if the file is missing."""
    try:
        path = next(os.path.join(ROOT, "profiles", f) for f in TRAFFIC_FILES if os.path.exists(os.path.join(ROOT, "profiles", f)))
        with open(path) as f:
            t = json.load(f)
        ks = ("gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2")
        return sum(t[k]["dram_read"] + t[k]["dram_write"] for k in ks) / len(ks)
    except Exception:
        return None


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m["hbm_gbs"]), bf16_tflops=float(m["bf16_tflops"]),
                 bf16_tflops_sustained=float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), source="measured")
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gflop_per_image(S, patch, dim, depth, heads, classes):
    """2*MAC, no padding, no recompute, vector-chain rollout excluded."""
    p = (S // patch) ** 2
    n = p + 1
    return (2 * p * 3 * patch * patch * dim + depth * (2 * n * dim * 3 * dim + 4 * heads * n * n * (dim // heads) + 2 * n * dim * dim
                                                         + 16 * n * dim * dim) + 4 * dim * classes + 2 * p * dim * classes) / 1e9


class Reference:
    """The reference's implementation of the path: the UNMODIFIED vit_model.py (oracle/_ref, copied by oracle/make_ref.py in the
    build container; imported through oracle/ref_shim.py) when it is there -> kind "reference"; otherwise the oracle port,
    bit-identical to it on this torch build (tests/golden/REPORT.json) -> kind "port".  forward + classic CAM (t.py:55-75 is
    not executable in the reference tree; oracle.postproc.classic_cam restates it for both kinds)."""

    def __init__(self, sd):
        from oracle import ref_shim, vit_forward as VF
        self.VF, self.shim, self.sd = VF, ref_shim, sd
        self.kind, self.model = "port", None
        if ref_shim.available():
            try:
                ref = ref_shim.import_reference()
                m = ref.vit_base_patch16_224_in21k(num_classes=C, has_logits=False)
                m.load_state_dict(sd, strict=True)
                m.eval()
                m.is_train = False
                self.model, self.kind = m, "reference"
            except Exception as e:      # noqa: BLE001 -- the port is the documented fallback of this (baseline-only) leg
                print(f"bench: reference import failed ({e!r}); timing the oracle port", file=sys.stderr)

    def to(self, device):
        import torch
        if self.model is not None:
            self.model = self.model.to(device)
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        self.device = torch.device(device)
        return self

    def step(self, x):
        """-> (logits [B,C], cam [B,C,14,14], hwp logits [B,C])"""
        import contextlib
        from oracle import postproc as PP
        if self.model is not None:
            ctx = self.shim.on_cpu() if not x.is_cuda else contextlib.nullcontext()
            with ctx:
                logits, attn_w, attn_m, hwp, w1, ori = self.model(x)
            return logits, PP.classic_cam(attn_m[-1], w1), hwp
        out = self.VF.forward(self.sd, x, self.VF.VIT_B16_224, keep_P=False)
        return out["logits"], PP.classic_cam(out["X"][-1], self.sd["head1.weight"]), out["hwp"]


def run_cpu(steps: int, warmup: int, batch: int, keep_outputs: bool = False):
    """The reference on the host cores (all of them), fp32: `warmup` untimed passes over 32 images, then `steps` timed passes
    over `batch` images (seeded synthetic images 0..batch-1, the reference's seed-0 weights)."""
    import torch
    from oracle import vit_forward as VF
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = Reference(VF.init_state_dict(VF.VIT_B16_224, 0))
    x = VF.make_images(0, batch)
    out = None
    with torch.no_grad():
        for _ in range(warmup):
            ref.step(x[:32])
        t0 = time.perf_counter()
        for _ in range(steps):
            out = ref.step(x)
        dt = time.perf_counter() - t0
    r = dict(value=steps * batch / dt, ms_per_step=dt / steps * 1e3, cores=cores, threads=torch.get_num_threads(), kind=ref.kind,
             sample=f"{steps} step(s) x {batch} images after {warmup} warm-up pass(es) over 32, fp32, forward + classic CAM "
                    f"(same weights / config as the GPU arm; {'reference vit_model.py' if ref.kind == 'reference' else 'oracle port'})")
    if keep_outputs:
        r["x"], r["out"], r["sd"] = x, out, ref.sd
    return r


def parity_block(model, dev, cpu):
    """GPU arm vs the CPU baseline's outputs on the same images and weights, at the north-star bars."""
    import numpy as np
    import torch
    from vision_transformer_cam_b200 import cam as CAM
    from oracle import postproc as PP
    x, (logits_ref, cam_ref, hwp_ref) = cpu["x"], cpu["out"]
    Bp = x.shape[0]
    o = model.forward_cam(x.to(dev))
    cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
    try:        # image-level label rows of the reference's voc12/cls_labels.npy (committed fixture)
        labels = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "default_b256.npz"))["labels"][:Bp]).float()
        src = "voc12/cls_labels.npy rows (tests/golden/default_b256.npz)"
    except Exception:
        labels = torch.zeros((Bp, C))
        labels[torch.arange(Bp), torch.arange(Bp) % C] = 1
        src = "synthetic one label per image"
    if labels.shape[0] < Bp:
        labels = labels.repeat((Bp + labels.shape[0] - 1) // labels.shape[0], 1)[:Bp]
    hw = (375, 500)
    lab = CAM.cam_pseudo_label(cam, labels.to(dev), hw).cpu()
    agree = []
    for i in range(0, Bp, 32):
        agree.append((lab[i:i + 32] == PP.cam_pseudo_label(cam_ref[i:i + 32], labels[i:i + 32], hw)).float().mean(dim=(1, 2)))
    agree = torch.cat(agree)
    cos = torch.nn.functional.cosine_similarity(cam.cpu().double().flatten(1), cam_ref.double().flatten(1), dim=1)
    rel = lambda a, b: float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())
    return {"images": Bp, "against": f"cpu_baseline outputs ({cpu['kind']}), same images and weights, fp32",
            "logits_max_rel_err": rel(o.logits, logits_ref), "cam_cosine_min": float(cos.min()), "cam_cosine_mean": float(cos.mean()),
            "pseudo_label_agreement": float(agree.mean()), "pseudo_label_agreement_min_image": float(agree.min()),
            "labels": src, "label_map_hw": list(hw),
            "bars": {"logits_max_rel_err": 1e-2, "cam_cosine": 0.999, "pseudo_label_agreement": 0.995},
            "pass": bool(rel(o.logits, logits_ref) <= 1e-2 and float(cos.min()) >= 0.999 and float(agree.mean()) >= 0.995)}


def gpu_eager_block(sd, dev, B, ours_ms):
    """The reference model itself, eager PyTorch on this GPU (cuBLAS / cuDNN / ATen), B images per step, torch.no_grad()."""
    import torch
    out = {"batch": B, "ours_ms_per_step": ours_ms}
    ref = Reference({k: v.clone() for k, v in sd.items()}).to(dev)
    out["kind"] = ref.kind
    x = torch.randn((B, 3, IMG, IMG), generator=torch.Generator(device=dev).manual_seed(1000), device=dev)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, ctx in (("fp32_tf32_off", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            def run():
                with torch.no_grad():
                    if ctx is None:
                        return ref.step(x)
                    with ctx:
                        return ref.step(x)
            run()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 3
            e0.record()
            for _ in range(n):
                run()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / n
            out[name] = {"ms_per_step": ms, "images_per_s": B / ms * 1e3, "speedup_of_ours": ms / ours_ms}
    except Exception as e:      # noqa: BLE001 -- an informational leg must not take the headline down
        out["error"] = repr(e)[:300]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    del ref, x
    torch.cuda.empty_cache()
    return out


def other_configs(dev, rank, world, pk):
    """BASELINE configs 3, 4, 5 (a few steps each; device-resident synthetic inputs, max over ranks)."""
    import torch
    import torch.distributed as dist
    import vision_transformer_cam_b200 as V
    from vision_transformer_cam_b200 import cam as CAM, pipeline as PIPE

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, iters):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / iters

    res = {}
    # config 3: 10,582 images (wc -l voc12/train_aug.txt), CAM + rollout, contiguous balanced shards, NCCL gather inside the timed region
    torch.manual_seed(0)
    model = V.vit_base_patch16_224_in21k(num_classes=C, has_logits=False).to(dev).eval()
    pool = torch.randn((256, 3, IMG, IMG), generator=torch.Generator(device=dev).manual_seed(2000 + rank), device=dev)
    n_items = 10582
    run3 = lambda n: PIPE.extract_cams_sharded(model, lambda lo, hi: pool[:hi - lo], n, batch=256, with_rollout=True, gather=True)
    run3(min(n_items, 512 * world))
    ms3 = timed(lambda: run3(n_items), 2)
    res["3"] = {"workload": "ViT-B/16 224px CAM + attention rollout over 10,582 synthetic images, batch-sharded, NCCL gather of cam / rollout / hwp logits",
                "images": n_items, "n_gpus": world, "scaling": "strong", "ms_per_pass": ms3, "images_per_s": n_items / ms3 * 1e3,
                "tflops_per_gpu": n_items / world * FLOP_IMAGE / (ms3 / 1e3) / 1e12}
    del model, pool
    torch.cuda.empty_cache()
    cases = (("4", "ViT-B/16 448px (785 tokens), forward + CAM + rollout through all 12 layers", dict(img_size=448, patch_size=16, embed_dim=768, depth=12, num_heads=12), True),
             ("5", "ViT-L/16 384px (577 tokens, 24 layers), forward + CAM, data-parallel", dict(img_size=384, patch_size=16, embed_dim=1024, depth=24, num_heads=16), False))
    Bc = 64
    for key, name, kw, rollout in cases:
        torch.manual_seed(0)
        model = V.VisionTransformer(num_classes=C, representation_size=None, **kw).to(dev).eval()
        x = torch.randn((Bc, 3, kw["img_size"], kw["img_size"]), generator=torch.Generator(device=dev).manual_seed(3000 + rank), device=dev)

        def step():
            o = model.forward_cam(x, rollout=rollout)
            cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
            return (cam, o.rollout) if rollout else (cam,)

        for _ in range(3):
            step()
        ms = timed(step, 6)
        gf = gflop_per_image(kw["img_size"], 16, kw["embed_dim"], kw["depth"], kw["num_heads"], C)
        tf = Bc * gf / ms
        res[key] = {"workload": name, "batch_per_gpu": Bc, "n_gpus": world, "scaling": "weak", "ms_per_step": ms,
                    "images_per_s": world * Bc / ms * 1e3, "gflop_per_image": round(gf, 3), "tflops_per_gpu": tf,
                    "frac_of_burst_peak": tf / pk["bf16_tflops"]}
        del model, x
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / eager-GPU-baseline / configs 3-5 legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "ViT-B/16 224px, 20-class head, batch 256 per GPU, forward + patch-token CAM (BASELINE configs[1])",
              "img": IMG, "batch_per_gpu": args.batch, "global_batch": args.batch * max(world, 1), "tokens": N,
              "gflop_per_image": round(FLOP_IMAGE / 1e9, 3), "mask_norm": "batch",
              "l2": "inputs larger than L2: 154 MB of images + ~1 GB of activations per step vs 126 MB L2",
              "collectives": ("none (1 GPU)" if world <= 1 else
                              "CAM maps all-gathered (NCCL) once per 8 steps on a side stream + at the end, int64 counters all-reduced at "
                              "the end; all inside the timed region, none inside the forward")}

    if args.impl == "reference":
        if rank != 0:
            return
        W = max(args.warmup, 1)
        r = run_cpu(max(args.steps, 1), W, args.batch)        # every timed step is the full batch the config names
        line = {"impl": "reference", "metric": "images/sec ViT-B/16 forward+CAM", "value": r["value"], "unit": "images/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import vision_transformer_cam_b200 as V
    from vision_transformer_cam_b200 import _lib, cam as CAM, pipeline as PIPE

    assert torch.cuda.is_available(), "bench.py needs a B200 (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its version banner there)
        if world == 2:      # two ranks: NCCL would spread the 4 MB-per-step gather over many channels; two CTAs move it and the
            os.environ.setdefault("NCCL_MAX_CTAS", "2")      # forward keeps 146 SMs.  (At 8 ranks the cap makes the collective itself slow.)
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    B = args.batch

    torch.manual_seed(0)
    model = V.vit_base_patch16_224_in21k(num_classes=C, has_logits=False).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x_dev = torch.randn((B, 3, IMG, IMG), generator=g, device=dev)
    GATHER_EVERY = 8       # steps per gather: every CAM is gathered inside the timed region
    counters = torch.zeros(4, dtype=torch.int64, device=dev)

    def local_step(x):
        o = model.forward_cam(x)
        cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
        return o, cam

    # The CAM gather: one NCCL all-gather per 8 steps on a side stream (dist.SideStreamGather; at 2 ranks NCCL is capped at two CTAs
    # so that it takes at most two SMs from the persistent kernels of the forward).  VTC_BENCH_GATHER=push selects peer-to-peer pushes by the
    # copy engines into symmetric memory instead (dist.PeerPushGather: no SM at all) -- measured at 2 GPUs: 11.15 ms per step
    # against 10.88 (NCCL) and 10.77 (NCCL, 2 CTAs); DESIGN.md section 6.
    from vision_transformer_cam_b200 import dist as VD
    gatherer = gathered = None
    gather_kind = "none (1 GPU)"
    if world > 1:
        want = os.environ.get("VTC_BENCH_GATHER", "nccl")
        if want == "push":
            try:
                gatherer = VD.PeerPushGather(dev, every=GATHER_EVERY)
                gathered = gatherer.alloc((world, GATHER_EVERY, B, C, 14, 14))
                gather_kind = ("CAM maps pushed into every peer's symmetric-memory buffer by device-to-device copies (copy engines over NVLink, "
                               "side stream) once per 8 steps + at the end")
            except Exception as e:      # noqa: BLE001
                print(f"bench: symmetric-memory gather unavailable ({e!r}); using the NCCL all-gather", file=sys.stderr)
                gatherer = None
        if gatherer is None:
            gatherer = VD.SideStreamGather(dev, every=GATHER_EVERY)
            gathered = torch.empty((world, GATHER_EVERY, B, C, 14, 14), device=dev)
            gather_kind = f"CAM maps all-gathered (NCCL, NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS', 'default')}) once per 8 steps on a side stream + at the end"
        config["collectives"] = gather_kind + "; int64 counters all-reduced (NCCL) at the end; all inside the timed region, none inside the forward"

    def step(x):
        o, cam = local_step(x)
        if world > 1:       # gather of the CAM maps + reduction of the counters: the only collectives (never inside the forward);
            gatherer.gather(gathered, cam)      # staged; one all-gather per GATHER_EVERY steps on a side stream, under the following forwards
        return o, cam

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        if world > 1:
            gatherer.wait()
            dist.all_reduce(counters)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident throughput
    for _ in range(W):
        step(x_dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.load().vtc_launch_count()
    t0 = time.time()
    ms = timed(lambda: step(x_dev), K)
    t1 = time.time()
    launches = _lib.load().vtc_launch_count() - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = world * B * K / (ms / 1e3)

    # ---- end to end through the public API: pinned host images -> H2D -> forward + CAM -> D2H of CAMs + logits.
    # Every step copies its own input batch from pinned host memory (pipeline.DeviceFeeder: the copy of step i+1 runs on a
    # side stream while step i computes) and reads its CAMs + logits back to the host.
    x_host = torch.randn((B, 3, IMG, IMG), generator=torch.Generator().manual_seed(1000 + rank)).pin_memory()
    cam_host = torch.empty((B, C, 14, 14)).pin_memory()
    logit_host = torch.empty((B, C)).pin_memory()
    feeder = PIPE.DeviceFeeder(dev)
    drain = PIPE.HostDrain(dev)

    def e2e_run(iters):
        for x in feeder.stream(x_host for _ in range(iters)):
            o, cam = step(x)
            drain.push(cam_host, cam)               # D2H on a side stream: the next step's kernels do not queue behind it
            drain.push(logit_host, o.logits)
        drain.wait()                                # the closing event below is recorded behind the last read-back

    e2e_run(2)
    # the host link of this box, measured alone (diagnostic: an e2e well below `value` on a box with a slow link is the link, not the path)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe = torch.empty_like(x_host, device=dev)
    probe.copy_(x_host, non_blocking=True)
    h0.record()
    for _ in range(3):
        probe.copy_(x_host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * x_host.numel() * 4 / (h0.elapsed_time(h1) * 1e-3) / 1e9
    del probe
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(K)
    if world > 1:
        gatherer.wait()
        dist.all_reduce(counters)
    e1.record()
    barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_t)
    e2e = {"value": world * B * K / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
           "d2h_bytes_per_step": (cam_host.numel() + logit_host.numel()) * 4, "ms_per_step": ms_e2e / K, "h2d_link_gbps_alone": round(h2d_gbps, 1),
           "overlap": "H2D of step i+1 on a copy stream during step i (pipeline.DeviceFeeder, two device buffers); D2H of step i on a second copy stream (pipeline.HostDrain), all of it inside the timed region"}

    # ---- dominant kernel, timed live with CUDA events around every launch of the same workload
    roof = None
    kernels = None
    if rank == 0:
        pk = peaks()
        model.kernel_profile(True)
        PK = 3
        for _ in range(PK):
            local_step(x_dev)           # rank 0 only: no collective here
        prof = model.kernel_profile()
        model.kernel_profile(False)
        gemm_ms = sum(prof[k][0] for k in ("gemm_patch", "gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2")) / PK
        gemm_n = sum(prof[k][1] for k in ("gemm_patch", "gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2")) // PK
        total_ms = sum(v[0] for v in prof.values()) / PK
        achieved = GEMM_FLOP_IMAGE * B / (gemm_ms / 1e3) / 1e12
        roof = {"kernel": "gemm_bf16_kernel (tcgen05, all epilogues)", "bound": "tensor", "achieved": achieved,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
                "peak_kind": f"bf16_tflops_sustained of {pk['source']} (kernel timed inside a long step); burst {pk['bf16_tflops']}",
                "frac_of_burst": achieved / pk["bf16_tflops"], "launches_per_step": gemm_n, "ms_per_step": gemm_ms,
                "share_of_step": gemm_ms / total_ms, "traffic": gemm_traffic(),
                "traffic_note": "mean DRAM bytes per GEMM launch (qkv, proj, fc1, fc2), one ncu --set full capture: profiles/r02_traffic.json (r01 if absent)",
                "whole_step_tflops": FLOP_IMAGE * B / (ms / K / 1e3) / 1e12,
                "whole_step_frac_of_burst": FLOP_IMAGE * B / (ms / K / 1e3) / 1e12 / pk["bf16_tflops"]}
        flops = {"gemm_patch": FLOP_PATCH, "gemm_qkv": L * FLOP_QKV, "gemm_proj": L * FLOP_PROJ, "gemm_fc1": L * FLOP_FC,
                 "gemm_fc2": L * FLOP_FC, "attention": L * FLOP_ATTN}
        kernels = {k: {"ms_per_step": round(v[0] / PK, 4), "launches": v[1] // PK,
                       **({"tflops": round(flops[k] * B / (v[0] / PK / 1e3) / 1e12, 1)} if k in flops and v[0] > 0 else {})}
                   for k, v in prof.items() if v[1] > 0}

    # ---- CPU baseline (rank 0, N = 1 only: about 10-15 s of host work) and, from its outputs, the parity block
    cpu = parity = eager = None
    if world == 1 and not args.no_cpu_baseline:
        r = run_cpu(2, 1, B, keep_outputs=True)
        cpu = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        if not args.no_extras:
            parity = parity_block(model, dev, r)
        sd_ref = r["sd"]
        del r
        if not args.no_extras:
            eager = gpu_eager_block(sd_ref, dev, B, ms / K)
    if world > 1:       # the last gathered block really holds every rank's CAMs (checked outside the timed region)
        torch.cuda.synchronize()
        dist.barrier()
        mine = gathered.view(world, -1)[rank].abs().sum()
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        seen = [float(gathered.view(world, -1)[r].abs().sum()) for r in range(world)]
        assert all(abs(seen[r] - float(sums[r])) <= 1e-3 * max(1.0, abs(seen[r])) for r in range(world)), (seen, [float(v) for v in sums])
    del x_dev, gathered, gatherer
    torch.cuda.empty_cache()
    configs = None if args.no_extras else other_configs(dev, rank, world, peaks())

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": "images/sec ViT-B/16 forward+CAM", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roof, "cpu_baseline": cpu, "parity": parity, "gpu_eager_baseline": eager, "configs": configs, "kernels": kernels,
                "tensor_frac_of_burst_peak": FLOP_IMAGE * value / world / 1e12 / peaks()["bf16_tflops"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
