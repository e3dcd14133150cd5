"""2 .. 8 GPUs (gpurun --gpus N): the sharded CAM extraction over NCCL equals the single-rank run; the side-stream NCCL gather
and the copy-engine peer-push gather deliver every rank's blocks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import dist as D, pipeline
from oracle import vit_forward as VF
rank, local_rank, world = D.init_from_env("nccl")
dev = torch.device("cuda", local_rank)
torch.manual_seed(0)
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
get = lambda lo, hi: VF.make_images(lo, hi - lo).to(dev)
n = 10
# batch 1 per call => the batch-global mask max is per image, so the result is independent of the sharding
out = pipeline.extract_cams_sharded(model, get, n_items=n, batch=1)
assert out["cam"].shape[0] == n and out["rollout"].shape == (n, 196)
if rank == 0:
    # the same per-image calls on one rank (same outputs requested => same kernels => same bits)
    from vision_transformer_cam_b200 import cam as CAM
    fwd = [model.forward_cam(get(i, i + 1), rollout=True) for i in range(n)]
    ref = torch.cat([CAM.classic_cam(o.tokens_last, model.head1.weight.data) for o in fwd])
    assert torch.equal(ref, out["cam"]), float((ref - out["cam"]).abs().max())
    assert torch.equal(torch.cat([o.rollout for o in fwd]), out["rollout"])
# side-stream gather: same result as the in-stream collective, for several steps in flight
g = D.SideStreamGather(dev)
outs = [torch.empty((world * 3, 5), device=dev) for _ in range(4)]
for k in range(4):
    local = torch.full((3, 5), float(10 * k + rank), device=dev)
    g.gather(outs[k], local)
    del local
g.wait()
for k in range(4):
    want = torch.cat([torch.full((3, 5), float(10 * k + r)) for r in range(world)])
    assert torch.equal(outs[k].cpu(), want)
# batched variant: 3 steps per collective, 7 steps => two full collectives + a flush of one step
g3 = D.SideStreamGather(dev, every=3)
out3 = torch.empty((world, 3, 3, 5), device=dev)
seen = []
for k in range(7):
    local = torch.full((3, 5), float(100 * k + rank), device=dev)
    g3.gather(out3, local)
    del local
    if k % 3 == 2:
        g3.wait()
        seen.append(out3.clone().cpu())
g3.wait()
for blk in range(2):
    for r in range(world):
        for j in range(3):
            assert torch.equal(seen[blk][r, j], torch.full((3, 5), float(100 * (3 * blk + j) + r)))
tail = out3.view(-1)[: world * 15].view(world, 1, 3, 5).cpu()
for r in range(world):
    assert torch.equal(tail[r, 0], torch.full((3, 5), float(600 + r)))
# copy-engine peer push into symmetric memory: 2 steps per push, 5 steps => two pushes + a flush of one step
pg = D.PeerPushGather(dev, every=2)
outp = pg.alloc((world, 2, 3, 5))
seenp = []
for k in range(5):
    local = torch.full((3, 5), float(1000 * k + rank), device=dev)
    pg.gather(outp, local)
    del local
    if k % 2 == 1:
        pg.wait()
        torch.cuda.synchronize()
        dist.barrier()                      # every peer's pushes have landed
        seenp.append(outp.clone().cpu())
        dist.barrier()                      # nobody overwrites before everybody has read
pg.wait()
torch.cuda.synchronize()
dist.barrier()
for blk in range(2):
    for r in range(world):
        for j in range(2):
            assert torch.equal(seenp[blk][r, j], torch.full((3, 5), float(1000 * (2 * blk + j) + r))), (blk, r, j)
tailp = outp.view(world, -1)[:, :15].cpu()
for r in range(world):
    assert torch.equal(tailp[r], torch.full((15,), float(4000 + r)))
c = torch.tensor([rank + 1], dtype=torch.int64, device=dev)
D.reduce_counters(c)
assert int(c) == world * (world + 1) // 2
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank %d ok\n" % rank); sys.stdout.flush()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_nccl_sharded_extraction_equals_single_rank(tmp_path, lib_built):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    nproc = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert all(f"rank {k} ok" in r.stdout for k in range(nproc)), r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_single_process_model_on_second_device(lib_built):
    """One process, two devices: a model living on cuda:1 runs while torch's current device is cuda:0 (the library keeps its
    shared-memory opt-in per device and every op launches on the stream of the device its tensors live on), and gives the bits
    of the same model on cuda:0."""
    import vision_transformer_cam_b200 as V
    from vision_transformer_cam_b200 import cam as CAM
    from oracle import vit_forward as VF
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    m0 = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).eval()
    sd = {k: v.clone() for k, v in m0.state_dict().items()}
    m0 = m0.to("cuda:0")
    torch.manual_seed(0)
    m1 = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).eval()
    m1.load_state_dict(sd)
    m1 = m1.to("cuda:1")
    x = VF.make_images(0, 3)
    a = m0.forward_cam(x.to("cuda:0"), rollout=True)
    assert torch.cuda.current_device() == 0
    b = m1.forward_cam(x.to("cuda:1"), rollout=True)              # first use of every kernel on device 1
    assert b.logits.device == torch.device("cuda:1") and torch.cuda.current_device() == 0
    assert torch.equal(a.logits.cpu(), b.logits.cpu()) and torch.equal(a.rollout.cpu(), b.rollout.cpu())
    ca = CAM.classic_cam(a.tokens_last, m0.head1.weight.data)
    cb = CAM.classic_cam(b.tokens_last, m1.head1.weight.data)     # ops.py wrappers on device 1 with device 0 current
    lab = torch.zeros((3, 20)); lab[:, 3] = 1
    la = CAM.cam_pseudo_label(ca, lab.to("cuda:0"), (64, 80))
    lb = CAM.cam_pseudo_label(cb, lab.to("cuda:1"), (64, 80))
    sa = CAM.hwp_pseudo_seg(a, m0.head1.weight.data, (64, 80))
    sb = CAM.hwp_pseudo_seg(b, m1.head1.weight.data, (64, 80))
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    assert torch.equal(ca.cpu(), cb.cpu()) and torch.equal(la.cpu(), lb.cpu()) and torch.equal(sa.cpu(), sb.cpu())
    with pytest.raises(RuntimeError):
        CAM.classic_cam(a.tokens_last, m1.head1.weight.data)      # tensors on different devices: loud
