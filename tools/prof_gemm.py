"""Time the forward's four GEMM shapes with the library VTC_LIB_PATH points to (default: the in-tree build): 8-launch bursts after
an idle gap.  For A/B runs of instrumented builds (tools/build_ablate.sh gN).    python tools/prof_gemm.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_transformer_cam_b200 import ops, _lib
dev = torch.device("cuda:0")
M = 50432
tag = os.path.basename(os.environ.get("VTC_LIB_PATH", "libvtc.so"))
for name, N, K, epi in (("qkv", 2304, 768, _lib.EPI_BIAS), ("fc1", 3072, 768, _lib.EPI_BIAS_GELU), ("fc2", 768, 3072, _lib.EPI_BIAS_RESIDUAL), ("proj", 768, 768, _lib.EPI_BIAS_RESIDUAL)):
    a = torch.randn((M, K), device=dev).bfloat16()
    w = (torch.randn((N, K), device=dev) * 0.02).bfloat16()
    b = torch.randn((N,), device=dev)
    res = torch.zeros((M, N), device=dev) if epi == _lib.EPI_BIAS_RESIDUAL else None
    out = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if res is None else res
    fn = (lambda: ops.gemm_bf16(a, w, b, epi, residual=res, out=out)) if res is not None else (lambda: ops.gemm_bf16(a, w, b, epi, out=out))
    fn(); fn()
    best = []
    for rep in range(3):
        time.sleep(0.3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / 8 * 1e3)
    print(f"{tag:28s} {name:5s} " + " ".join(f"{u:7.1f} us" for u in best) + f"   {2.0 * M * N * K / min(best) / 1e6:7.1f} TF/s", flush=True)
