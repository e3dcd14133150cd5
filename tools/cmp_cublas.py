"""GEMM kernel vs cuBLAS (torch) on the forward's own shapes, short bursts with idle gaps (no power cap) and a long run
(power cap).  Debug / measurement tool; prints TFLOP/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_transformer_cam_b200 import ops, _lib
dev = torch.device("cuda:0")
M = 50432
shapes = [("qkv", 2304, 768, _lib.EPI_BIAS), ("fc1", 3072, 768, _lib.EPI_BIAS_GELU), ("fc2", 768, 3072, _lib.EPI_BIAS_RESIDUAL), ("proj", 768, 768, _lib.EPI_BIAS_RESIDUAL)]


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, N, K, epi in shapes:
    a = torch.randn((M, K), device=dev).bfloat16()
    w = (torch.randn((N, K), device=dev) * 0.02).bfloat16()
    b = torch.randn((N,), device=dev)
    bb = b.bfloat16()
    res = torch.zeros((M, N), device=dev) if epi == _lib.EPI_BIAS_RESIDUAL else None
    out = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if res is None else res
    mine = (lambda: ops.gemm_bf16(a, w, b, epi, residual=res, out=out)) if res is not None else (lambda: ops.gemm_bf16(a, w, b, epi, out=out))
    cub = lambda: torch.nn.functional.linear(a, w, bb)
    fl = 2.0 * M * N * K
    for label, fn in (("vtc", mine), ("cublas", cub)):
        fn(); fn()
        time.sleep(0.5)
        burst = timed(fn, 8)
        time.sleep(0.5)
        long_ = timed(fn, 400)
        print(f"{name:5s} {label:7s} burst {fl / burst / 1e9:7.1f} TF/s ({burst * 1e3:6.1f} us)   400 back-to-back {fl / long_ / 1e9:7.1f} TF/s", flush=True)
