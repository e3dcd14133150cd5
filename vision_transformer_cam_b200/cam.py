"""CAM / attention-rollout / pseudo-label extraction: the math that the reference runs inline in predict.py and
validate.py after the forward, as callable functions over the libvtc kernels.

The reference has no API for this layer (it is script code), so this module defines one; every function cites the
reference lines whose result it reproduces.  Inputs are `CamForward` fields (vit_model.forward_cam) or the entries of
the reference 6-tuple; everything stays on the GPU."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .vit_model import CamForward


REF_KEPT_LAYERS = 12


def last_kept(per_layer: torch.Tensor) -> torch.Tensor:
    """The reference's forward only returns the last 12 layers' attention / tokens (vit_model.py:322, `len(blocks) - i <=
    12`), and predict.py / validate.py index into THAT list: for depth 24 'layers[5:]' are blocks 17..23 and the rollout
    runs over blocks 12..23.  [L,...] -> [min(L,12),...] (a contiguous view)."""
    return per_layer[-REF_KEPT_LAYERS:] if per_layer.shape[0] > REF_KEPT_LAYERS else per_layer


def patch_similarity(tokens: torch.Tensor) -> torch.Tensor:
    """predict.py:191-199 (viz): `F.normalize(x).squeeze(0) @ ....t()` for block outputs x [B,N,D]; F.normalize's default
    dim=1 normalises every feature column across the tokens (reproduced as is).  -> [B,N,N]."""
    return ops.patch_similarity(tokens.float().contiguous())


# ---- attention rollout (predict.py:189-247) --------------------------------------------------------------------
def head_mean(attn_weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """predict.py:189-190 for a batch: list of L [B,H,N,N] -> [L,B,N,N]."""
    return torch.stack([ops.head_mean(p.contiguous()) for p in attn_weights])


def rollout_row(attn_mean: torch.Tensor) -> torch.Tensor:
    """predict.py:215-232: CLS row of prod_l (mean_l + I)/rowsum, patch columns, un-normalised.  [L,B,N,N] -> [B,P]."""
    return ops.rollout(last_kept(attn_mean).contiguous())


def rollout_map(attn_mean: torch.Tensor, out_hw: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """predict.py:229-247: rollout row / max, g x g, bilinear (cv2.resize == align_corners=False) to (H,W).  Takes the head
    means [L,B,N,N] or an already computed rollout row [B,P] (`forward_cam(rollout=True).rollout`)."""
    row = ops.normalize_max_(rollout_row(attn_mean) if attn_mean.dim() == 4 else attn_mean.clone())
    B, P = row.shape
    g = int(round(P ** 0.5))
    m = row.view(B, g, g)
    return m if out_hw is None else ops.upsample_bilinear(m, out_hw)


def layer_maps(cls_rows: torch.Tensor, out_hw: Optional[Tuple[int, int]] = None, as_u8: bool = False) -> torch.Tensor:
    """predict.py:261-269: per layer (mean_h P + I)/rowsum, CLS row, / max -> [min(L,12),B,g,g] (or resized, optionally *255 u8)."""
    cls_rows = last_kept(cls_rows).contiguous()
    L, B, H, N = cls_rows.shape
    g = int(round((N - 1) ** 0.5))
    m = torch.stack([ops.cls_layer_map(cls_rows, l, l + 1) for l in range(L)]).view(L, B, g, g)
    if out_hw is None:
        return m
    return ops.upsample_bilinear(m, out_hw, as_u8=as_u8)


# ---- classic CAM (t.py:55-75, utils.py:80-129) -------------------------------------------------------------------
def classic_cam(tokens_last: torch.Tensor, head1_weight: torch.Tensor, relu: bool = True, eps: float = 1e-5) -> torch.Tensor:
    """cam[b,c] = minmax(relu(F_b . W_c)) on the block-L patch tokens -> [B,C,g,g] in [0,1]."""
    return ops.cam_project(tokens_last.contiguous(), head1_weight.contiguous(), relu, eps)


def cam_pseudo_label(cam: torch.Tensor, labels: torch.Tensor, out_hw: Tuple[int, int], bg_thresh: float = 0.25) -> torch.Tensor:
    """utils.py:100-108 convention: upsample, keep the image's classes, argmax against a constant background score.
    Fused upsample+argmax: writes 1 byte per pixel.  -> uint8 [B,H,W] (0 = background, c+1)."""
    return ops.cam_label(cam.contiguous(), labels, out_hw, bg_thresh)


def cam_upsampled(cam: torch.Tensor, out_hw: Tuple[int, int], as_u8: bool = False) -> torch.Tensor:
    """Full-resolution maps on request (utils.py:86,113: uint8(255*cam) resized)."""
    return ops.upsample_bilinear(cam.contiguous(), out_hw, as_u8=as_u8)


# ---- validate.py:132-258 pseudo segmentation ---------------------------------------------------------------------
def bg_map(cls_rows: torch.Tensor, first_layer: int = 5) -> torch.Tensor:
    """validate.py:225-237: mean CLS attention over layers[first_layer:] of the kept (last 12) layers and heads, + identity,
    renormalised, / max."""
    cls_rows = last_kept(cls_rows).contiguous()
    return ops.cls_layer_map(cls_rows, first_layer, cls_rows.shape[0])


def hwp_pseudo_seg(fwd: CamForward, head1_weight: torch.Tensor, out_hw: Tuple[int, int], sig_thresh: float = 0.9,
                   cos_thresh: float = 0.5, bg_thresh: float = 0.05, return_parts: bool = False):
    """validate.py:132-258 for a batch -> uint8 [B,H,W].  'Patch owns no feature' (the reference's sentinel >= 21, which
    overflows its confusion matrix) is mapped to background."""
    p2c, cos = ops.hwp_cos_vote(fwd.hwp_logits, head1_weight.contiguous(), fwd.hwp_tokens, fwd.tokens_last.contiguous(), sig_thresh)
    bgm = bg_map(fwd.cls_rows)
    seg = ops.hwp_seg(cos, p2c, bgm, out_hw, cos_thresh, bg_thresh)
    return (seg, p2c, bgm) if return_parts else seg


# ---- metrics (utils.py:30-77, 248-262) -----------------------------------------------------------------------------
class ConfusionMatrix:
    """utils.py:30-77 with the counters on the GPU (int64 [n,n], n = num_classes + 1)."""

    def __init__(self, num_classes: int, device="cuda"):
        self.num_classes = num_classes
        self.n = num_classes + 1
        self.mat = torch.zeros((self.n, self.n), dtype=torch.int64, device=device)

    def update(self, gt: torch.Tensor, pred: torch.Tensor) -> None:
        """gt / pred: uint8 label maps of equal size; gt >= n (255) is ignored (utils.py:42)."""
        ops.confmat_update(self.mat, gt.to(torch.uint8).contiguous().view(-1), pred.to(torch.uint8).contiguous().view(-1))

    def reset(self) -> None:
        self.mat.zero_()

    def compute(self):
        h = self.mat.float()
        acc_global = torch.diag(h).sum() / h.sum()
        acc = torch.diag(h) / h.sum(1)
        iu = torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))
        return acc_global, acc, iu

    def reduce_from_all_processes(self) -> None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.mat)

    def __str__(self):
        acc_global, acc, iu = self.compute()
        return ("global correct: {:.1f}\naverage row correct: {}\nIoU: {}\nmean IoU: {:.1f}").format(
            acc_global.item() * 100, ["{:.1f}".format(i) for i in (acc * 100).tolist()],
            ["{:.1f}".format(i) for i in (iu * 100).tolist()], iu.nanmean().item() * 100)


def average_precision(labels: torch.Tensor, scores: torch.Tensor, acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sklearn.metrics.average_precision_score per image (what utils.py:258 calls on the host), computed on the GPU:
    labels / scores [B,C] -> ap [B] fp64, -1 for images without a positive label.  `acc` (fp64 [2] on the device)
    accumulates (sum of APs, number of scored images) without any device->host traffic."""
    return ops.average_precision(labels.float().contiguous(), scores.float().contiguous(), acc)


def compute_mAP(labels: torch.Tensor, outputs: torch.Tensor) -> List[float]:
    """utils.py:248-262: per-image AP over the class scores, for images with at least one positive label."""
    ap = average_precision(labels.to(outputs.device), outputs)
    return [a for a in ap.tolist() if a >= 0.0]
