"""The callers either side of the fused forward, as library functions: what predict.py:129-293 computes for an image
(rollout map, per-layer CLS maps, class scores) and what validate.py:121-292 computes over a dataset (pseudo
segmentation, confusion matrix / mIoU, mAP) -- batched, on the GPU, optionally sharded over the GPUs of one box.

File / image IO, plotting and argparse of the reference drivers are out of scope; inputs here are already
normalised image tensors [B,3,S,S] (the output of the reference's transforms, predict.py:72-75)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Iterator, Optional, Tuple

import torch

from . import cam as CAM
from . import dist as D
from .vit_model import VisionTransformer


class DeviceFeeder:
    """Host -> device ingest overlapped with compute: batch i+1 is copied from (pinned) host memory on a side stream into
    the second of two device buffers while batch i runs through the forward on the caller's stream.

        for x in DeviceFeeder(device).stream(host_batches):     # x: device tensor, valid until the next iteration
            out = model.forward_cam(x)

    Batches that already live on the device are passed through untouched."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs: list = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]      # copy into buffer k finished
        self.free = [torch.cuda.Event() for _ in range(depth)]       # compute on buffer k finished

    def _start_copy(self, k: int, host: torch.Tensor, used_before: bool) -> torch.Tensor:
        buf = self.bufs[k]
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = self.bufs[k] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            # the caching allocator may hand out a block whose previous owner still has kernels queued on the compute stream:
            # reuse is only ordered on THAT stream, so the first copy into a new buffer waits for it
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            if used_before:
                self.copy_stream.wait_event(self.free[k])
            buf.copy_(host, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        return buf

    def stream(self, batches: Iterable[torch.Tensor]) -> Iterator[torch.Tensor]:
        it = iter(batches)
        compute = torch.cuda.current_stream(self.device)
        pending = []                                   # [(slot or None, tensor)] copies in flight, oldest first
        n = 0

        def launch() -> bool:
            nonlocal n
            try:
                b = next(it)
            except StopIteration:
                return False
            if b.is_cuda:
                pending.append((None, b))
            else:
                k = n % self.depth
                pending.append((k, self._start_copy(k, b, n >= self.depth)))
                n += 1
            return True

        for _ in range(self.depth - 1):
            launch()
        while True:
            launch()                                   # keep `depth - 1` copies ahead of the batch being consumed
            if not pending:
                return
            k, x = pending.pop(0)
            if k is not None:
                compute.wait_event(self.ready[k])
            yield x
            if k is not None:
                self.free[k].record(compute)


class HostDrain:
    """Device -> host read-back overlapped with compute, the counterpart of `DeviceFeeder`: a step's results are copied
    device-to-device into a small staging ring on the caller's stream (a few microseconds) and from there into (pinned) host
    tensors on a side stream, so the next step's kernels do not queue behind the PCIe transfer.

        drain = HostDrain(device)
        for x in batches:
            o = model.forward_cam(x)
            drain.push(cam_host, cam)          # returns at once; cam may be dropped by the caller
        drain.wait()                           # the caller's stream (and the host, if sync=True) sees every copy finished

    The staging ring (not `Tensor.record_stream`) is what keeps the device tensor alive: a block the caching allocator cannot
    hand out again until a side-stream event has passed makes the next step's allocations fall through to cudaMalloc, which
    synchronises the device.  A host tensor handed to `push` must not be read before `wait()`."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.rings: dict = {}              # (shape, dtype) -> [buffers, staged events, drained events, pushes so far]
        self.done = torch.cuda.Event()

    def push(self, host: torch.Tensor, dev: torch.Tensor) -> None:
        if tuple(host.shape) != tuple(dev.shape) or host.dtype != dev.dtype or host.is_cuda or not dev.is_cuda:
            raise ValueError(f"HostDrain.push: host {tuple(host.shape)} {host.dtype} / device {tuple(dev.shape)} {dev.dtype} do not match")
        key = (tuple(dev.shape), dev.dtype)
        ring = self.rings.get(key)
        if ring is None:
            ring = self.rings[key] = [[torch.empty_like(dev) for _ in range(self.depth)], [torch.cuda.Event() for _ in range(self.depth)],
                                      [torch.cuda.Event() for _ in range(self.depth)], 0]
        bufs, staged, drained, n = ring
        k = n % self.depth
        ring[3] = n + 1
        compute = torch.cuda.current_stream(self.device)
        if n >= self.depth:
            compute.wait_event(drained[k])         # the read-back that used this staging buffer two pushes ago
        bufs[k].copy_(dev)                         # device to device, on the caller's stream
        staged[k].record(compute)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(staged[k])
            host.copy_(bufs[k], non_blocking=True)
            drained[k].record(self.copy_stream)
            self.done.record(self.copy_stream)

    def wait(self, sync: bool = False) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)
        if sync:
            self.done.synchronize()


@dataclass
class Prediction:
    """predict.py outputs for a batch."""
    logits: torch.Tensor            # [B,C] head
    hwp_scores: torch.Tensor        # [B,C] sigmoid(hwp logits)         predict.py:291
    rollout: torch.Tensor           # [B,H,W] rollout map / max          predict.py:229-247
    layer_maps: torch.Tensor        # [L,B,H,W] uint8 per-layer CLS maps predict.py:261-269
    cam: torch.Tensor               # [B,C,g,g] classic CAM              t.py:55-75


@torch.no_grad()
def predict(model: VisionTransformer, x: torch.Tensor, out_hw: Optional[Tuple[int, int]] = None) -> Prediction:
    """predict.py:129-293 for a batch of normalised images (the reference runs batch 1)."""
    out_hw = out_hw or (x.shape[-2], x.shape[-1])
    o = model.forward_cam(x, rollout=True)
    return Prediction(logits=o.logits, hwp_scores=torch.sigmoid(o.hwp_logits), rollout=CAM.rollout_map(o.rollout, out_hw),
                      layer_maps=CAM.layer_maps(o.cls_rows, out_hw, as_u8=True),
                      cam=CAM.classic_cam(o.tokens_last, model.head1.weight.data))


class Validator:
    """validate.py:118-292: pseudo segmentation + confusion matrix + per-image AP, accumulated on the GPU and reduced over
    ranks at the end.  `mask_norm='image'` makes every image independent of its batch (the reference validates at batch 1,
    where the batch-global max of vit_model.py:335 is per image)."""

    def __init__(self, model: VisionTransformer, num_classes: int = 20, device="cuda"):
        self.model = model
        self.num_classes = num_classes
        self.confmat = CAM.ConfusionMatrix(num_classes, device=device)
        self.device = torch.device(device)
        self.ap_acc = torch.zeros(2, dtype=torch.float64, device=self.device)     # (sum of APs, scored images), on the GPU

    @torch.no_grad()
    def step(self, x: torch.Tensor, target: Optional[torch.Tensor] = None, seg_labels: Optional[torch.Tensor] = None,
             out_hw: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """x [B,3,S,S]; target [B,C] multi-hot (mAP, validate.py:266-273); seg_labels [B,H,W] uint8 with 255 = ignore
        (confusion matrix, validate.py:276).  Returns the pseudo segmentation uint8 [B,H,W]."""
        out_hw = out_hw or ((seg_labels.shape[-2], seg_labels.shape[-1]) if seg_labels is not None else (x.shape[-2], x.shape[-1]))
        o = self.model.forward_cam(x, mask_norm="image")
        seg = CAM.hwp_pseudo_seg(o, self.model.head1.weight.data, out_hw)
        if seg_labels is not None:
            self.confmat.update(seg_labels.to(self.device), seg)
        if target is not None:      # per-image AP on the device (validate.py:266-273 does it on the host through sklearn)
            CAM.average_precision(target.to(self.device), torch.sigmoid(o.hwp_logits), self.ap_acc)
        return seg

    def finalize(self) -> Dict[str, object]:
        """All-reduce the counters (NCCL / gloo) and compute global accuracy, per-class IoU, mIoU and mAP."""
        stats = self.ap_acc.clone()
        D.reduce_counters(self.confmat.mat)
        D.reduce_counters(stats)
        acc_global, acc, iu = self.confmat.compute()
        return dict(global_acc=float(acc_global), per_class_acc=acc.tolist(), iou=iu.tolist(), miou=float(iu.nanmean()),
                    mAP=float(stats[0] / stats[1]) if float(stats[1]) > 0 else float("nan"), confmat=self.confmat.mat.clone())


@torch.no_grad()
def extract_cams_sharded(model: VisionTransformer, get_images: Callable[[int, int], torch.Tensor], n_items: int, batch: int = 256,
                         with_rollout: bool = True, gather: bool = True) -> Dict[str, torch.Tensor]:
    """BASELINE config 3: CAM (+ rollout) extraction over an `n_items`-image set, batch-sharded over the ranks of the
    process group.  `get_images(lo, hi)` returns the normalised images [hi-lo,3,S,S] of global indices [lo,hi), either on
    this rank's GPU or in (pinned) host memory -- host batches are copied in on a side stream while the previous batch
    computes (`DeviceFeeder`).  Returns (on every rank when `gather`) cam [n,C,g,g], rollout [n,g*g] and hwp_logits [n,C]
    in global image order; collectives: one all_gather per output, outside the forward."""
    rank, world = D.rank_world()
    lo, hi = D.shard_range(n_items, rank, world)
    cams, rolls, hwps = [], [], []
    device = next(model.parameters()).device
    for x in DeviceFeeder(device).stream(get_images(b0, b1) for b0, b1 in D.batches(lo, hi, batch)):
        o = model.forward_cam(x, rollout=with_rollout)
        cams.append(CAM.classic_cam(o.tokens_last, model.head1.weight.data))
        hwps.append(o.hwp_logits)
        if with_rollout:
            rolls.append(o.rollout)
    out = {"cam": torch.cat(cams), "hwp_logits": torch.cat(hwps)}
    if with_rollout:
        out["rollout"] = torch.cat(rolls)
    if gather and world > 1:
        out = {k: D.gather_shards(v, n_items) for k, v in out.items()}
    return out
