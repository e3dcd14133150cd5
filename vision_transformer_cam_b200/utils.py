"""The callers of the forward in the reference's utils.py, on top of the fused path:

  * `evaluate` -- utils.py:205-244: mAP of the 196-patch head and of the 16-patch (high-weight-patch) head over a loader;
  * `teacher_cams` -- the no-grad pass that produces class activation maps while training (`update_log.md:5`, the former
    `cams` model output consumed by utils.py:80-129): classic CAM, the per-image "syn" map (max over the image's classes,
    utils.py:118-120) and the label-restricted pseudo label map, all computed on the GPU;
  * `ConfusionMatrix`, `compute_mAP` -- utils.py:30-77, 248-262 (re-exported from `cam`).

`train_one_epoch` (utils.py:144-203) needs a backward pass and stays out of scope."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import cam as CAM
from . import dist as D
from .cam import ConfusionMatrix, compute_mAP  # noqa: F401
from .vit_model import VisionTransformer


def _forward_any(model: VisionTransformer, image: torch.Tensor, device, **kw):
    """Normalised fp32 NCHW (the reference's loader) or decoded uint8 HWC (voc12.make_u8_loader) batches."""
    image = image.to(device, non_blocking=True)
    if image.dtype == torch.uint8:
        return model.forward_cam_u8(image, **kw)
    return model.forward_cam(image, **kw)


@torch.no_grad()
def evaluate(model: VisionTransformer, data_loader, device, epoch: int = 0, num_classes: int = 20, reduce: bool = False) -> Tuple[float, float]:
    """utils.py:205-244.  Loader items are (name, image, target[, seg_labels]); returns (mAP of sigmoid(head logits), mAP of
    sigmoid(high-weight-patch logits)) over the images with at least one positive label.  The per-image average precision
    (sklearn on the host in the reference, utils.py:258) is accumulated on the device; nothing is copied back until the end.
    Like the reference the result is per rank unless `reduce` (then the two counters are all-reduced first)."""
    del epoch, num_classes                          # the reference only uses them for its progress-bar text
    model.eval()
    model.is_train = False
    device = torch.device(device)
    acc196 = torch.zeros(2, dtype=torch.float64, device=device)
    acc16 = torch.zeros(2, dtype=torch.float64, device=device)
    for data in data_loader:
        image, target = data[1], data[2].to(device, non_blocking=True)
        o = _forward_any(model, image, device)
        CAM.average_precision(target, torch.sigmoid(o.logits), acc196)
        CAM.average_precision(target, torch.sigmoid(o.hwp_logits), acc16)
    if reduce:
        D.reduce_counters(acc196)
        D.reduce_counters(acc16)
    a, b = acc196.tolist(), acc16.tolist()
    return (a[0] / a[1] if a[1] > 0 else float("nan")), (b[0] / b[1] if b[1] > 0 else float("nan"))


@torch.no_grad()
def teacher_cams(model: VisionTransformer, images: torch.Tensor, labels: Optional[torch.Tensor] = None,
                 out_hw: Optional[Tuple[int, int]] = None, bg_thresh: float = 0.25) -> Dict[str, torch.Tensor]:
    """CAM generation during training: one fused no-grad forward of the current weights (the module's parameters are read
    as they are; packed copies are refreshed when their versions changed) ->
        logits [B,C], hwp_logits [B,C], cam [B,C,g,g] in [0,1];
        with `labels` [B,C] multi-hot: syn_cam [B,g,g] = max over the image's classes (utils.py:118-120) and, with `out_hw`,
        pseudo_label uint8 [B,H,W] (0 = background, c + 1), the fused upsample + argmax of cam.cam_pseudo_label.
    The training mode flags of the module are left untouched."""
    device = next(model.parameters()).device
    o = _forward_any(model, images, device)
    cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
    out = {"logits": o.logits, "hwp_logits": o.hwp_logits, "cam": cam}
    if labels is not None:
        labels = labels.to(device)
        keep = (labels > 0).to(cam.dtype)[:, :, None, None]
        out["syn_cam"] = (cam * keep).amax(dim=1)
        if out_hw is not None:
            out["pseudo_label"] = CAM.cam_pseudo_label(cam, labels.float().contiguous(), out_hw, bg_thresh)
    return out
