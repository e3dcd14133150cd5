"""CPU: the C-ABI library loads and exports exactly what include/vtc.h declares; the drop-in module surface; sharding
logic; world-size-2 gloo collectives.  No compute call is made (there is no GPU here and no CPU path in the product)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vtc.h")).read()
    return sorted(set(re.findall(r"VTC_API\s+[\w\s\*]+?\b(vtc_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    from vision_transformer_cam_b200 import _lib
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", lib_built], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (vtc_\w+)", out)))
    assert set(declared) <= set(exported), set(declared) - set(exported)
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert lib.vtc_version() == 100
    for name in declared:
        assert getattr(lib, name) is not None


def test_library_is_sm100a_only_and_has_tcgen05(lib_built):
    sass = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG"):        # tcgen05.mma / tcgen05.ld / TMA load, store, reduce
        assert mnemonic in sass, mnemonic
    # no legacy mma.sync tensor path for the dense contractions of the forward: HMMA only inside the general-shape attention
    # kernel that serves head dimensions other than 64 (ViT-H/14, csrc/attention_generic.cu) and the HBM-bound [P x D] x [D x 20]
    # CAM projection (split-bf16 operands, csrc/postproc.cu), where the warp-level form is the right size
    func = ""
    for line in sass.splitlines():
        if "Function :" in line:
            func = line
        elif " HMMA." in line:                      # (UTCHMMA is the tcgen05 mnemonic)
            assert "attention_generic" in func or "cam_project" in func, func


def test_no_gpu_means_loud_failure(lib_built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vision_transformer_cam_b200 import _lib
    lib = _lib.load()
    assert lib.vtc_check_device() == -4                                        # VTC_ERR_CUDA: no device, no fallback
    assert b"no CPU path" in lib.vtc_last_error() or b"CUDA error" in lib.vtc_last_error()
    import vision_transformer_cam_b200 as V
    m = V.VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2, num_classes=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32))


def test_model_create_rejects_unsupported_shapes(lib_built):
    import ctypes
    from vision_transformer_cam_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    ok = _lib.Config(224, 16, 3, 20, 768, 12, 12, 3072, 0, 4, 0.25, 16, 1e-6)
    assert lib.vtc_model_create(ctypes.byref(ok), ctypes.byref(h)) == 0
    assert lib.vtc_model_packed_bytes(h) == 2 * (768 * 768 + 12 * (3 * 768 * 768 + 768 * 768 + 2 * 3072 * 768))
    assert lib.vtc_workspace_bytes(h, 256, None) > 0
    assert lib.vtc_model_destroy(h) == 0
    vith = _lib.Config(224, 14, 3, 20, 1280, 32, 16, 5120, 0, 4, 0.25, 16, 1e-6)   # ViT-H/14: patch 14 (K 588 -> 640), head_dim 80
    assert lib.vtc_model_create(ctypes.byref(vith), ctypes.byref(h)) == 0
    assert lib.vtc_model_packed_bytes(h) == 2 * (1280 * 640 + 32 * (3 * 1280 * 1280 + 1280 * 1280 + 2 * 5120 * 1280))
    assert lib.vtc_model_destroy(h) == 0
    bad = _lib.Config(224, 16, 3, 20, 768, 12, 16, 3072, 0, 4, 0.25, 16, 1e-6)      # head_dim 48 runs; 768 / 32 = 24 does not
    assert lib.vtc_model_create(ctypes.byref(bad), ctypes.byref(h)) == 0 and lib.vtc_model_destroy(h) == 0
    bad = _lib.Config(224, 16, 3, 20, 768, 12, 32, 3072, 0, 4, 0.25, 16, 1e-6)
    assert lib.vtc_model_create(ctypes.byref(bad), ctypes.byref(h)) == -2 and b"head_dim" in lib.vtc_last_error()
    bad = _lib.Config(448, 16, 3, 20, 1280, 12, 16, 5120, 0, 4, 0.25, 16, 1e-6)     # head_dim 80 with 785 tokens: general-shape kernel limit
    assert lib.vtc_model_create(ctypes.byref(bad), ctypes.byref(h)) == -2 and b"tokens" in lib.vtc_last_error()


def test_dropin_module_surface_and_state_dict():
    import vision_transformer_cam_b200.vit_model as VM
    for name in ("VisionTransformer", "Block", "Attention", "Mlp", "PatchEmbed", "DropPath", "drop_path", "_init_vit_weights",
                 "vit_base_patch16_224", "vit_base_patch16_224_in21k", "vit_base_patch32_224", "vit_base_patch32_224_in21k",
                 "vit_large_patch16_224", "vit_large_patch16_224_in21k", "vit_large_patch32_224_in21k", "vit_huge_patch14_224_in21k"):
        assert hasattr(VM, name), name
    from oracle import vit_forward as VF
    torch.manual_seed(0)
    m = VM.vit_base_patch16_224_in21k(num_classes=20, has_logits=False)
    ref = VF.init_state_dict(VF.VIT_B16_224, 0)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys()) or set(sd) == set(ref)
    assert max(float((sd[k] - ref[k]).abs().max()) for k in sd) == 0.0          # same RNG consumption as the reference
    assert m.has_logits is False and m.is_train is True and m.head1.weight.shape == (20, 768)
    m2 = VM.vit_base_patch16_224_in21k(num_classes=5, has_logits=True)
    assert "pre_logits.fc.weight" in m2.state_dict() and m2.has_logits
    missing, unexpected = m2.load_state_dict({k: v for k, v in m2.state_dict().items() if not k.startswith("head.")}, strict=False)
    assert set(missing) == {"head.weight", "head.bias"} and not unexpected      # the reference's checkpoint-loading idiom
    import copy, pickle
    copy.deepcopy(m2)
    pickle.loads(pickle.dumps(m2))


def test_unsupported_constructor_arguments_are_rejected_not_ignored():
    """The fused forward computes scale = head_dim ** -0.5, erf-GELU and LayerNorm; a constructor argument it cannot honour
    must raise instead of being silently dropped (the module-level Attention would honour it, the fused path would not)."""
    import torch.nn as nn
    from functools import partial
    import vision_transformer_cam_b200 as V
    kw = dict(img_size=32, patch_size=16, embed_dim=256, depth=1, num_heads=4, num_classes=3)
    V.VisionTransformer(**kw)                                              # fine
    V.VisionTransformer(qk_scale=64 ** -0.5, **kw)                         # the default value spelled out: fine
    V.VisionTransformer(norm_layer=partial(nn.LayerNorm, eps=1e-5), **kw)  # any eps: fine (passed to the kernels)
    for bad in (dict(qk_scale=0.3), dict(act_layer=nn.ReLU), dict(norm_layer=nn.BatchNorm1d), dict(qkv_bias=False), dict(distilled=True),
                dict(drop_ratio=0.1), dict(norm_layer=partial(nn.LayerNorm, elementwise_affine=False))):
        with pytest.raises(NotImplementedError):
            V.VisionTransformer(**bad, **kw)


def test_shard_ranges_and_batches():
    from vision_transformer_cam_b200 import dist as D
    assert D.shard_sizes(10582, 8) == [1323] * 6 + [1322] * 2
    covered = []
    for r in range(8):
        lo, hi = D.shard_range(10582, r, 8)
        covered += list(range(lo, hi))
    assert covered == list(range(10582))
    b = D.batches(0, 1323, 256)
    assert len(b) == 6 and {e - s for s, e in b} == {220, 221} and b[0][0] == 0 and b[-1][1] == 1323
    assert D.batches(5, 5, 256) == [] and D.shard_range(3, 7, 8) == (3, 3)


WORKER = r'''
import os, sys, torch
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from vision_transformer_cam_b200 import dist as D
rank, local_rank, world = D.init_from_env("gloo")
assert world == 2 and D.rank_world() == (rank, 2)
n = 11
lo, hi = D.shard_range(n, rank, world)
local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3) * 10
full = D.gather_shards(local, n)
assert torch.equal(full, torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3) * 10), full
c = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int64)
D.reduce_counters(c)
assert c.tolist() == [3, 30]
assert D.max_over_ranks(float(rank), "cpu") == 1.0
dist.destroy_process_group()
sys.stdout.write("rank %d ok\n" % rank); sys.stdout.flush()
'''


def test_gloo_world_size_2_gather_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket
    with socket.socket() as sock:            # a free port: a fixed one collides with a run that has just finished
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own vit_model.py from oracle/_ref when the recipe oracle/make_ref.py has
    run, else the CPU oracle port) prints one JSON line with the contract's keys, at the batch it was asked for."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "8"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["config"]["batch_per_gpu"] == 8 and "1 step(s) x 8 images" in line["cpu_baseline"]["sample"]
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "vit_model.py")) or os.path.exists("/root/reference/vit_model.py")
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_copy_recipe_and_shim():
    """oracle/make_ref.py writes a byte-identical, git-ignored copy of the reference model; the shim finds it."""
    import hashlib
    import json
    from oracle import make_ref, ref_shim
    if not os.path.exists("/root/reference/vit_model.py"):
        pytest.skip("build container only")
    assert make_ref.make()
    man = json.load(open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")))
    assert man["sha256"]["vit_model.py"] == hashlib.sha256(open("/root/reference/vit_model.py", "rb").read()).hexdigest()
    assert ref_shim.available()
    ignored = subprocess.run(["git", "-C", ROOT, "check-ignore", "oracle/_ref/vit_model.py"], capture_output=True, text=True)
    assert ignored.returncode == 0            # never part of the history
