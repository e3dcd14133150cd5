"""Oracle (test infrastructure): CAM / rollout / pseudo-label post-processing, restated.

Each function cites the reference lines it follows.  Inputs are the outputs of
`oracle.vit_forward.forward` (or of the reference model itself); everything is fp32 torch on
CPU, integer label maps are uint8 / int64 as in the reference.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# predict.py:189-247  attention rollout; predict.py:261-269 per-layer CLS maps
# ------------------------------------------------------------------------------------------------
def head_mean(P_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """predict.py:189-190 generalised to a batch: [L,B,N,N] head-mean attention."""
    return torch.stack([P.mean(dim=1) for P in P_list])


def augment(att: torch.Tensor) -> torch.Tensor:
    """predict.py:215-218: add the identity (residual path) and renormalise rows."""
    eye = torch.eye(att.size(-1), dtype=att.dtype)
    aug = att + eye
    return aug / aug.sum(dim=-1).unsqueeze(-1)


def rollout_dense(P_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """predict.py:221-232: joint[0]=A0; joint[n]=A_n @ joint[n-1]; CLS row, patch columns.
    Returns the un-normalised [B, N-1] rollout row (the reference then divides by its max)."""
    aug = augment(head_mean(P_list))                      # [L,B,N,N]
    joint = aug[0]
    for n in range(1, aug.size(0)):
        joint = torch.matmul(aug[n], joint)
    return joint[:, 0, 1:]


def rollout_map(P_list: Sequence[torch.Tensor], out_hw: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """predict.py:229-247: reshape to g x g, divide by max, bilinear resize to (H,W).
    cv2.resize(INTER_LINEAR) on float == half-pixel bilinear == align_corners=False
    (probe 1e-6), so F.interpolate is used for both."""
    r = rollout_dense(P_list)
    B, P = r.shape
    g = int(round(P ** 0.5))
    m = (r / r.max(dim=1, keepdim=True).values).reshape(B, 1, g, g)
    if out_hw is None:
        return m[:, 0]
    return F.interpolate(m, size=out_hw, mode="bilinear", align_corners=False)[:, 0]


def layer_maps(P_list: Sequence[torch.Tensor], out_hw: Optional[Tuple[int, int]] = None,
               as_u8: bool = False) -> torch.Tensor:
    """predict.py:261-269: per layer aug[l][0,1:] / max -> g x g (-> resize -> *255 -> uint8).
    Returns [L,B,g,g] (or [L,B,H,W])."""
    aug = augment(head_mean(P_list))
    rows = aug[:, :, 0, 1:]                               # [L,B,P]
    L, B, P = rows.shape
    g = int(round(P ** 0.5))
    m = (rows / rows.max(dim=2, keepdim=True).values).reshape(L * B, 1, g, g)
    if out_hw is not None:
        m = F.interpolate(m, size=out_hw, mode="bilinear", align_corners=False)
    m = m.reshape(L, B, *m.shape[-2:])
    if as_u8:
        return (m * 255).to(torch.uint8)                  # .astype("uint8") truncation, predict.py:269
    return m


# ------------------------------------------------------------------------------------------------
# classic CAM: t.py:55-75, utils.py:80-88 (cam_norm), vit_model.py:297 (ReLU), SURVEY.md A.3
# ------------------------------------------------------------------------------------------------
def classic_cam(X_last: torch.Tensor, head1_weight: torch.Tensor, relu: bool = True,
                eps: float = 1e-5) -> torch.Tensor:
    """cam[b,c,p] = sum_d W[c,d] F[b,p,d] (t.py:66) on the block-L patch tokens (the space head1 is
    trained in, vit_model.py:381-393); ReLU; per (b,c) map: subtract min, divide by max
    (utils.py:84-85; + eps so an all-zero map gives 0 instead of the reference's 0/0).
    Returns [B,C,g,g] fp32 in [0,1]."""
    Fp = X_last[:, 1:, :]
    B, P, D = Fp.shape
    g = int(round(P ** 0.5))
    cam = torch.einsum("bpd,cd->bcp", Fp, head1_weight)
    if relu:
        cam = torch.relu(cam)
    cam = cam - cam.min(dim=2, keepdim=True).values
    cam = cam / (cam.max(dim=2, keepdim=True).values + eps)
    return cam.reshape(B, -1, g, g)


def cam_pseudo_label(cam: torch.Tensor, labels: torch.Tensor, out_hw: Tuple[int, int],
                     bg_thresh: float = 0.25) -> torch.Tensor:
    """WSSS convention (utils.py:100-108 keeps only the image's classes; argmax against a constant
    background score).  cam [B,C,g,g], labels [B,C] in {0,1}.  Returns uint8 [B,H,W], 0 = bg, c+1."""
    up = F.interpolate(cam, size=out_hw, mode="bilinear", align_corners=False)   # [B,C,H,W]
    up = torch.where(labels[:, :, None, None] > 0, up, torch.full_like(up, -1.0))
    bg = torch.full((cam.shape[0], 1, *out_hw), bg_thresh, dtype=up.dtype)
    return torch.cat([bg, up], dim=1).argmax(dim=1).to(torch.uint8)


# ------------------------------------------------------------------------------------------------
# validate.py:132-208  high-weight-patch pseudo segmentation (vectorised form)
# ------------------------------------------------------------------------------------------------
def hwp_patch_classes(hwp_logits: torch.Tensor, head1_weight: torch.Tensor, ori: torch.Tensor,
                      sig_thresh: float = 0.9) -> torch.Tensor:
    """validate.py:132-153 for one image.  hwp_logits [C], head1_weight [C,D], ori [K,D].
    Returns patch_to_cls [K] int64 (class index 0..C-1; a sentinel >= 21 when a patch owns no
    feature, exactly the reference's arange(21, ...) table entries)."""
    C, D = head1_weight.shape
    K = ori.shape[0]
    pred = torch.sigmoid(hwp_logits) >= sig_thresh                         # :132-134
    W = head1_weight.clone()
    W[~pred] = -10.0                                                       # :138-142
    cls_of_feat = torch.argmax(W, dim=0)                                   # :143  [D]
    owner = torch.argmax(ori, dim=0)                                       # :148  [D]
    table = torch.arange(21, 21 + D * K, 1).reshape(D, K)                  # :146
    table[torch.arange(D), owner] = cls_of_feat                            # :150-151
    patch_to_cls, _ = torch.mode(table, dim=0)                             # :153
    return patch_to_cls


def hwp_cos_maps(X_last_img: torch.Tensor, ori_img: torch.Tensor) -> torch.Tensor:
    """validate.py:163-175: cosine similarity of the K hw-patch tokens with all patch tokens.
    X_last_img [N,D], ori_img [K,D] -> [K,g,g]."""
    patch = X_last_img[1:, :]
    g = int(round(patch.shape[0] ** 0.5))
    a = F.normalize(patch, dim=1)
    b = F.normalize(ori_img, dim=1)
    return (b @ a.t()).reshape(-1, g, g)


def bg_map(cls_rows: torch.Tensor, first_layer: int = 5) -> torch.Tensor:
    """validate.py:225-237: mean over layers[first_layer:] and heads of the CLS attention row, add the
    identity, renormalise, drop the CLS column, divide by the max.  cls_rows [L,B,H,N] -> [B,g*g].
    (Only the CLS row of the head/layer-mean matrix is used, so CLS rows suffice.)"""
    row = cls_rows[first_layer:].mean(dim=0).mean(dim=1)       # [B,N]; reference: mean(layers) then mean(heads)
    row = row.clone()
    row[:, 0] = row[:, 0] + 1.0                                # + identity on the CLS row
    row = row / row.sum(dim=1, keepdim=True)
    m = row[:, 1:]
    return m / m.max(dim=1, keepdim=True).values


def hwp_pseudo_seg(hwp_logits: torch.Tensor, head1_weight: torch.Tensor, ori: torch.Tensor,
                   X_last: torch.Tensor, cls_rows: torch.Tensor, out_hw: Tuple[int, int],
                   cos_thresh: float = 0.5, bg_thresh: float = 0.05,
                   clamp_sentinel: bool = False) -> torch.Tensor:
    """validate.py:132-258 for a batch (the reference runs batch 1).  Returns uint8 [B,H,W].

    clamp_sentinel=False reproduces the reference bit-for-bit (labels >= 22 wrap through uint8);
    True maps the 'patch owns no feature' sentinel to background (SURVEY.md appendix B decision,
    what the CUDA path does)."""
    B = hwp_logits.shape[0]
    h, w = out_hw
    bgm = bg_map(cls_rows)
    g = int(round(bgm.shape[1] ** 0.5))
    outs = []
    for b in range(B):
        p2c = hwp_patch_classes(hwp_logits[b], head1_weight, ori[b])
        cos = hwp_cos_maps(X_last[b], ori[b])                                       # [K,g,g]
        up = F.interpolate(cos.unsqueeze(0), size=(h, w), mode="bilinear", align_corners=False)[0]  # :177
        kstar = up.argmax(dim=0)                                                    # :179
        vmax = up.max(dim=0).values                                                 # :180
        fg = (vmax >= cos_thresh).to(torch.float32)                                 # :183-186
        seg = p2c[kstar] + 1                                                        # :190-208
        if clamp_sentinel:
            seg = torch.where(seg > 21, torch.zeros_like(seg), seg)
        bgu = F.interpolate(bgm[b].reshape(1, 1, g, g), size=(h, w), mode="bilinear",
                            align_corners=False)[0, 0]                              # :239-241
        bgk = (bgu >= bg_thresh).to(torch.float32)                                  # :244-246
        res = seg.to(torch.float32) * (fg * bgk)                                    # :248-257
        outs.append(res.to(torch.uint8))                                            # :258
    return torch.stack(outs)


# ------------------------------------------------------------------------------------------------
# utils.py:30-77 ConfusionMatrix; utils.py:248-262 compute_mAP
# ------------------------------------------------------------------------------------------------
def confmat_update(mat: Optional[np.ndarray], gt: np.ndarray, pred: np.ndarray, num_classes: int = 20) -> np.ndarray:
    """utils.py:35-45: n = C+1; keep GT in [0,n); bincount(n*gt + pred).  int64 [n,n].
    Predictions >= n (the reference's sentinel overflow) would index out of range in the
    reference's bincount reshape; here they are counted nowhere (callers clamp first)."""
    n = num_classes + 1
    if mat is None:
        mat = np.zeros((n, n), dtype=np.int64)
    gt = gt.reshape(-1).astype(np.int64)
    pred = pred.reshape(-1).astype(np.int64)
    k = (gt >= 0) & (gt < n) & (pred < n)
    inds = n * gt[k] + pred[k]
    mat += np.bincount(inds, minlength=n * n).reshape(n, n)
    return mat


def confmat_compute(mat: np.ndarray):
    """utils.py:51-61: global accuracy, per-class accuracy, per-class IoU (float32 like the reference)."""
    h = mat.astype(np.float32)
    diag = np.diag(h)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc_global = diag.sum() / h.sum()
        acc = diag / h.sum(1)
        iu = diag / (h.sum(1) + h.sum(0) - diag)
    return acc_global, acc, iu


def average_precision(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """sklearn.metrics.average_precision_score (binary) restated: AP = sum_n (R_n - R_{n-1}) P_n over
    distinct score thresholds in decreasing order (utils.py:258 calls it per image)."""
    y_true = np.asarray(y_true).astype(np.float64)
    y_score = np.asarray(y_score).astype(np.float64)
    order = np.argsort(-y_score, kind="mergesort")
    y_true, y_score = y_true[order], y_score[order]
    distinct = np.where(np.diff(y_score))[0]
    thr_idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true)[thr_idx]
    fps = 1 + thr_idx - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def compute_mAP(labels: np.ndarray, scores: np.ndarray) -> List[float]:
    """utils.py:248-262: per-image AP for images with at least one positive label."""
    out = []
    for i in range(labels.shape[0]):
        if labels[i].sum() > 0:
            out.append(average_precision(labels[i], scores[i]))
    return out
