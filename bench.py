#!/usr/bin/env python
"""Headline benchmark: images/s of the ViT-B/16 224 px forward + patch-token CAM (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of B synthetic images per GPU: fused forward (patch embed, 12
blocks with the background mask, hwp head, final norm + head) + classic CAM projection (20 x 14 x 14 maps per image).
Prints ONE JSON line (rank 0).  `value` is whole-job images/s with the inputs resident in HBM; `e2e` the same metric with
pinned-host inputs copied H2D and the CAMs + logits copied D2H inside the timed region; `roofline` the dominant kernel
(the tcgen05 GEMM) timed live with CUDA events; `cpu_baseline` the oracle port of the reference on the host cores.
Under torchrun (N > 1) every rank processes its own batch (weak scaling); the per-step CAMs are all-gathered and the
counters all-reduced over NCCL inside the timed region; time = max over ranks.
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


def gemm_traffic():
    """DRAM bytes per GEMM launch (launch-weighted mean over the qkv / proj / fc1 / fc2 GEMMs of a layer) from the committed
    ncu --set full capture, profiles/r01_traffic.json; None if the file is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
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


def cpu_reference_step(sd, cfg, x):
    """The reference's CPU path for this workload, via the oracle port (bit-identical to the reference on the same torch
    build, tests/golden/REPORT.json): forward (vit_model.py:303-424) + classic CAM."""
    from oracle import vit_forward as VF, postproc as PP
    out = VF.forward(sd, x, cfg, keep_P=False)
    cam = PP.classic_cam(out["X"][-1], sd["head1.weight"])
    return out["logits"], cam


def run_cpu(steps: int, warmup: int, batch: int):
    import torch
    from oracle import vit_forward as VF
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = VF.VIT_B16_224
    sd = VF.init_state_dict(cfg, 0)
    x = VF.make_images(0, batch)
    with torch.no_grad():
        for _ in range(warmup):
            cpu_reference_step(sd, cfg, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(sd, cfg, x)
        dt = time.perf_counter() - t0
    return dict(value=steps * batch / dt, ms_per_step=dt / steps * 1e3, cores=cores, threads=torch.get_num_threads(),
                sample=f"{steps} steps x {batch} images, fp32, forward + classic CAM (same weights/config as the GPU arm)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        r = run_cpu(max(args.steps, 1), W, 16)
        line = {"impl": "reference", "metric": "images/sec ViT-B/16 forward+CAM", "value": r["value"], "unit": "images/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
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
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    B = args.batch

    torch.manual_seed(0)
    model = V.vit_base_patch16_224_in21k(num_classes=C, has_logits=False).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x_dev = torch.randn((B, 3, IMG, IMG), generator=g, device=dev)
    GATHER_EVERY = 8       # steps per collective (dist.SideStreamGather): every CAM is gathered inside the timed region
    gathered = torch.empty((world, GATHER_EVERY, B, C, 14, 14), device=dev) if world > 1 else None
    counters = torch.zeros(4, dtype=torch.int64, device=dev)

    def local_step(x):
        o = model.forward_cam(x)
        cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
        return o, cam

    from vision_transformer_cam_b200 import dist as VD
    gatherer = VD.SideStreamGather(dev, every=GATHER_EVERY) if world > 1 else None

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

    def e2e_run(iters):
        for x in feeder.stream(x_host for _ in range(iters)):
            o, cam = step(x)
            cam_host.copy_(cam, non_blocking=True)
            logit_host.copy_(o.logits, non_blocking=True)

    e2e_run(2)
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
           "d2h_bytes_per_step": (cam_host.numel() + logit_host.numel()) * 4, "ms_per_step": ms_e2e / K,
           "overlap": "H2D of step i+1 on a copy stream during step i (pipeline.DeviceFeeder, two device buffers)"}

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
                "traffic_note": "mean DRAM bytes per GEMM launch (qkv, proj, fc1, fc2), one ncu --set full capture: profiles/r01_traffic.json",
                "whole_step_tflops": FLOP_IMAGE * B / (ms / K / 1e3) / 1e12,
                "whole_step_frac_of_burst": FLOP_IMAGE * B / (ms / K / 1e3) / 1e12 / pk["bf16_tflops"]}
        flops = {"gemm_patch": FLOP_PATCH, "gemm_qkv": L * FLOP_QKV, "gemm_proj": L * FLOP_PROJ, "gemm_fc1": L * FLOP_FC,
                 "gemm_fc2": L * FLOP_FC, "attention": L * FLOP_ATTN}
        kernels = {k: {"ms_per_step": round(v[0] / PK, 4), "launches": v[1] // PK,
                       **({"tflops": round(flops[k] * B / (v[0] / PK / 1e3) / 1e12, 1)} if k in flops and v[0] > 0 else {})}
                   for k, v in prof.items() if v[1] > 0}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        r = run_cpu(4, 1, 16)
        cpu = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": "images/sec ViT-B/16 forward+CAM", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
                "tensor_frac_of_burst_peak": FLOP_IMAGE * value / world / 1e12 / peaks()["bf16_tflops"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
