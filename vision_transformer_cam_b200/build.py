"""Build libvtc.so (hand-written sm_100a CUDA behind a C-ABI) in-tree with plain nvcc.

    python -m vision_transformer_cam_b200.build [--force]

nvcc cross-compiles without a GPU.  `-gencode arch=compute_100a,code=sm_100a` is required (the plain
`-arch=sm_100a` form also emits a generic compute_100 PTX pass that ptxas rejects for tcgen05).
The library links cudart statically and has no link-time libcuda dependency, so it loads on a CPU-only host.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libvtc.so")
SOURCES = ["runtime.cu", "gemm.cu", "elementwise.cu", "attention.cu", "attention_kv.cu", "attention_cs.cu", "attention_generic.cu", "cls_ops.cu", "postproc.cu", "model.cu"]
HEADERS = [os.path.join(CSRC, h) for h in ("common.cuh", "ops.h", "tma_host.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "vtc.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if _stale(obj, [path] + HEADERS):
        cmd = [NVCC] + NVCC_FLAGS + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
        if os.path.exists(LIB_PATH):
            os.remove(LIB_PATH)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if _stale(LIB_PATH, objs):
        cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print("built", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
