"""Tensor-level wrappers over the C-ABI kernels (one function per `vtc_*` entry point).

PyTorch is plumbing here: it owns device memory and the stream; every computation is a libvtc kernel.  All
wrappers validate device / dtype / contiguity and raise (never fall back) when something is off."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


class _DevPtr(int):
    """A device pointer that remembers the device it lives on (ctypes takes it as a plain integer)."""
    dev: torch.device


def _ptr(t: Optional[torch.Tensor], dtype=None, name: str = "tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (libvtc has no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    p = _DevPtr(t.data_ptr())
    p.dev = t.device
    return p


def _call(name: str, *args) -> None:
    """Enqueue `vtc_<name>(*args, stream)` on the current stream of the device the tensors live on (not torch's current
    device: a model on cuda:1 must launch on cuda:1), with that device made current for the launch."""
    dev = None
    for a in args:
        if isinstance(a, _DevPtr):
            if dev is None:
                dev = a.dev
            elif a.dev != dev:
                raise RuntimeError(f"{name}: tensors on different devices ({dev} and {a.dev})")
    if dev is None:
        raise RuntimeError(f"{name}: no device tensor among the arguments")
    with torch.cuda.device(dev):
        _lib.call(name, *args, torch.cuda.current_stream(dev).cuda_stream)


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _call("vtc_cast_bf16", _ptr(src, torch.float32, "src"), _ptr(out), src.numel())
    return out


def gemm_bf16(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, epilogue: int = _lib.EPI_BIAS,
              residual: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, tokens: int = 0) -> torch.Tensor:
    """a [M,K] bf16, w [N,K] bf16 (nn.Linear layout), bias [N] fp32."""
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        if epilogue in (_lib.EPI_BIAS, _lib.EPI_BIAS_GELU):
            out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        elif epilogue == _lib.EPI_BIAS_RESIDUAL:
            out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        else:
            raise RuntimeError("patch-embed epilogue needs an explicit token buffer `out`")
    _call("vtc_gemm_bf16", _ptr(a, torch.bfloat16, "a"), _ptr(w, torch.bfloat16, "w"), _ptr(bias, torch.float32, "bias"),
              _ptr(residual, torch.float32, "residual"), _ptr(pos, torch.float32, "pos"), _ptr(out), M, N, K, epilogue, tokens)
    return out


def split_bf16(src: torch.Tensor) -> torch.Tensor:
    """fp32 [..., K] -> (hi | lo) bf16 halves [..., 2K] with src ~= hi + lo (the operand format of the fp32 mode)."""
    K = src.shape[-1]
    rows = src.numel() // K
    out = torch.empty(src.shape[:-1] + (2 * K,), dtype=torch.bfloat16, device=src.device)
    _call("vtc_split_bf16", _ptr(src, torch.float32, "src"), _ptr(out), rows, K)
    return out


def merge_split(x: torch.Tensor) -> torch.Tensor:
    """(hi | lo) halves [..., 2K] -> fp32 [..., K] (host-side helper for tests)."""
    K = x.shape[-1] // 2
    return x[..., :K].float() + x[..., K:].float()


def gemm_split(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, epilogue: int = _lib.EPI_BIAS,
               residual: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, tokens: int = 0) -> torch.Tensor:
    """fp32-mode GEMM: a [M,2K], w [N,2K] split operands; bf16 outputs come back as split halves [M,2N]."""
    M, K2 = a.shape
    K = K2 // 2
    N = w.shape[0]
    assert w.shape[1] == K2
    if out is None:
        if epilogue in (_lib.EPI_BIAS, _lib.EPI_BIAS_GELU):
            out = torch.empty((M, 2 * N), dtype=torch.bfloat16, device=a.device)
        elif epilogue == _lib.EPI_BIAS_RESIDUAL:
            out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        else:
            raise RuntimeError("patch-embed epilogue needs an explicit token buffer `out`")
    _call("vtc_gemm_split", _ptr(a, torch.bfloat16, "a"), _ptr(w, torch.bfloat16, "w"), _ptr(bias, torch.float32, "bias"),
              _ptr(residual, torch.float32, "residual"), _ptr(pos, torch.float32, "pos"), _ptr(out), M, N, K, epilogue, tokens)
    return out


def patchify(x: torch.Tensor, patch: int, split: bool = False) -> torch.Tensor:
    B, Cin, S, S2 = x.shape
    assert S == S2
    g = S // patch
    out = torch.empty((B * g * g, Cin * patch * patch * (2 if split else 1)), dtype=torch.bfloat16, device=x.device)
    _call("vtc_patchify_split" if split else "vtc_patchify", _ptr(x, torch.float32, "x"), _ptr(out), B, Cin, S, patch)
    return out


def cls_token_rows(cls_token: torch.Tensor, pos_embed: torch.Tensor, tokens: torch.Tensor) -> None:
    B, N, D = tokens.shape
    _call("vtc_cls_token_rows", _ptr(cls_token, torch.float32), _ptr(pos_embed, torch.float32), _ptr(tokens, torch.float32),
              B, N, D)


def layernorm_bf16(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float, split: bool = False) -> torch.Tensor:
    D = x.shape[-1]
    rows = x.numel() // D
    out = torch.empty(x.shape[:-1] + (D * (2 if split else 1),), dtype=torch.bfloat16, device=x.device)
    _call("vtc_layernorm_split" if split else "vtc_layernorm_bf16", _ptr(x, torch.float32, "x"), _ptr(weight, torch.float32),
              _ptr(bias, torch.float32), _ptr(out), rows, D, eps)
    return out


def attention(qkv: torch.Tensor, heads: int, scale: float, key_bias: Optional[torch.Tensor] = None, want_cls: bool = True,
              want_attn: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """qkv [B,N,3*H*64] bf16 -> (out [B,N,H*64] bf16, cls_rows [B,H,N] fp32 | None, attn [B,H,N,N] fp32 | None)."""
    B, N, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((B, N, D), dtype=torch.bfloat16, device=qkv.device)
    cls = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device) if want_cls else None
    attn = torch.empty((B, heads, N, N), dtype=torch.float32, device=qkv.device) if want_attn else None
    _call("vtc_attention", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(key_bias, torch.float32, "key_bias"), _ptr(out), _ptr(cls),
              _ptr(attn), B, N, heads, scale)
    return out, cls, attn


def attention_mean(qkv: torch.Tensor, heads: int, scale: float, key_bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Attention that also returns the head mean of P without a [B,H,N,N] round trip: (out, cls_rows, attn_mean [B,N,N])."""
    B, N, D3 = qkv.shape
    out = torch.empty((B, N, D3 // 3), dtype=torch.bfloat16, device=qkv.device)
    cls = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device)
    mean = torch.empty((B, N, N), dtype=torch.float32, device=qkv.device)
    nbytes = int(_lib.load().vtc_attention_mean_scratch_bytes(B, N, heads))
    scratch = torch.empty((nbytes,), dtype=torch.uint8, device=qkv.device)
    _call("vtc_attention_mean", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(key_bias, torch.float32, "key_bias"), _ptr(out), _ptr(cls), _ptr(mean),
              _ptr(scratch), nbytes, B, N, heads, scale)
    return out, cls, mean


def attention_generic(qkv: torch.Tensor, heads: int, scale: float, key_bias: Optional[torch.Tensor] = None, want_cls: bool = True,
                      want_attn: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """The general-shape attention path (any head_dim that is a multiple of 16 up to 128, N <= 320): (out, cls_rows, attn)."""
    B, N, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((B, N, D), dtype=torch.bfloat16, device=qkv.device)
    cls = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device) if want_cls else None
    attn = torch.empty((B, heads, N, N), dtype=torch.float32, device=qkv.device) if want_attn else None
    _call("vtc_attention_generic", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(key_bias, torch.float32, "key_bias"), _ptr(out), _ptr(cls), _ptr(attn),
              B, N, heads, D // heads, scale)
    return out, cls, attn


def attention_kv(qkv: torch.Tensor, heads: int, scale: float, key_bias: Optional[torch.Tensor] = None, want_cls: bool = True,
                 want_attn: bool = False, split: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """The KV-blocked kernel called directly (any N <= 2048).  split: qkv [B,N,2*3*H*64] and out [B,N,2*H*64] hold
    (hi | lo) bf16 halves per token (`split_bf16`)."""
    B, N, W = qkv.shape
    D = W // (6 if split else 3)
    out = torch.empty((B, N, D * (2 if split else 1)), dtype=torch.bfloat16, device=qkv.device)
    cls = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device) if want_cls else None
    attn = torch.empty((B, heads, N, N), dtype=torch.float32, device=qkv.device) if want_attn else None
    _call("vtc_attention_kv", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(key_bias, torch.float32, "key_bias"), _ptr(out), _ptr(cls),
              _ptr(attn), B, N, heads, scale, int(split))
    return out, cls, attn


def head_mean(attn: torch.Tensor) -> torch.Tensor:
    B, H, N, _ = attn.shape
    out = torch.empty((B, N, N), dtype=torch.float32, device=attn.device)
    _call("vtc_head_mean", _ptr(attn, torch.float32), _ptr(out), B, H, N)
    return out


def cls_stat(cls_rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    B, H, N = cls_rows.shape
    cmap = torch.empty((B, N - 1), dtype=torch.float32, device=cls_rows.device)
    gmax = torch.zeros((1,), dtype=torch.float32, device=cls_rows.device)
    _call("vtc_cls_stat", _ptr(cls_rows, torch.float32), _ptr(cmap), _ptr(gmax), B, H, N)
    return cmap, gmax


def cls_mask(cls_map: torch.Tensor, gmax: torch.Tensor, thresh: float = 0.25, per_image: bool = False,
             forced_bg: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    B, P = cls_map.shape
    bg = torch.empty((B, P), dtype=torch.uint8, device=cls_map.device)
    kb = torch.empty((B, P + 1), dtype=torch.float32, device=cls_map.device)
    _call("vtc_cls_mask", _ptr(cls_map, torch.float32), _ptr(gmax, torch.float32), _ptr(forced_bg, torch.uint8), thresh,
              int(per_image), _ptr(bg), _ptr(kb), B, P + 1)
    return bg, kb


def cls_stat_mask(cls_rows: torch.Tensor, thresh: float = 0.25, per_image: bool = False, forced_bg: Optional[torch.Tensor] = None,
                  scale: Optional[float] = None):
    """cls_stat + cls_mask in one launch (the forward's kernel): -> (cls_map [B,P], gmax [1], bg [B,P] u8, key_bias [B,N], mask
    operands uint8 [B, bytes] of the fast attention kernel when `scale` (the attention scale) is given, else None)."""
    B, H, N = cls_rows.shape
    dev = cls_rows.device
    cmap = torch.empty((B, N - 1), dtype=torch.float32, device=dev)
    gmax = torch.zeros((1,), dtype=torch.float32, device=dev)
    bg = torch.empty((B, N - 1), dtype=torch.uint8, device=dev)
    kb = torch.empty((B, N), dtype=torch.float32, device=dev)
    ticket = torch.zeros((1,), dtype=torch.int32, device=dev)
    aug = None
    if scale is not None:
        aug = torch.zeros((B, int(_lib.load().vtc_attention_mask_operand_bytes(N))), dtype=torch.uint8, device=dev)
    _call("vtc_cls_stat_mask", _ptr(cls_rows, torch.float32, "cls_rows"), _ptr(cmap), _ptr(gmax), _ptr(forced_bg, torch.uint8), thresh, int(per_image),
          _ptr(bg), _ptr(kb), _ptr(ticket), _ptr(aug), (1.0 / scale) if scale is not None else 0.0, B, H, N)
    return cmap, gmax, bg, kb, aug


def attention_masked(qkv: torch.Tensor, heads: int, scale: float, key_bias: torch.Tensor, mask_operands: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fast attention with the precomputed mask operands of `cls_stat_mask` (what the forward does): (out, cls_rows)."""
    B, N, D3 = qkv.shape
    out = torch.empty((B, N, D3 // 3), dtype=torch.bfloat16, device=qkv.device)
    cls = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device)
    _call("vtc_attention_masked", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(key_bias, torch.float32, "key_bias"), _ptr(mask_operands, torch.uint8, "mask_operands"),
          _ptr(out), _ptr(cls), B, N, heads, scale)
    return out, cls


def rollout(attn_mean: torch.Tensor) -> torch.Tensor:
    """attn_mean [L,B,N,N] fp32 -> un-normalised rollout row [B,N-1] (predict.py:215-232)."""
    L, B, N, _ = attn_mean.shape
    out = torch.empty((B, N - 1), dtype=torch.float32, device=attn_mean.device)
    _call("vtc_rollout", _ptr(attn_mean, torch.float32), _ptr(out), L, B, N)
    return out


def rollout_operand_ld(n_tokens: int) -> int:
    return int(_lib.load().vtc_rollout_operand_ld(n_tokens))


def rollout_operand_from_mean(attn_mean: torch.Tensor) -> torch.Tensor:
    """fp32 head means [..., B, N, N] -> bf16 rollout operands [..., B, N, ldr] (row = N values | zero padding | fp32 row sum)."""
    N = attn_mean.shape[-1]
    rows = attn_mean.numel() // (N * N)
    out = torch.empty(attn_mean.shape[:-1] + (rollout_operand_ld(N),), dtype=torch.bfloat16, device=attn_mean.device)
    _call("vtc_rollout_operand_from_mean", _ptr(attn_mean, torch.float32, "attn_mean"), _ptr(out), rows, N)
    return out


def rollout_operands(operands: torch.Tensor, n_tokens: int) -> torch.Tensor:
    """bf16 rollout operands [L,B,N,ldr] -> un-normalised rollout row [B,N-1] (the streaming kernel of the fused forward)."""
    L, B, N, ldr = operands.shape
    assert N == n_tokens and ldr == rollout_operand_ld(N)
    out = torch.empty((B, N - 1), dtype=torch.float32, device=operands.device)
    _call("vtc_rollout_operands", _ptr(operands, torch.bfloat16, "operands"), _ptr(out), L, B, N)
    return out


def attention_mean_operand(qkv: torch.Tensor, heads: int, scale: float, key_bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Attention that leaves the head mean of P as a bf16 rollout operand: (out, cls_rows, operand [B,N,ldr])."""
    B, N, D3 = qkv.shape
    out = torch.empty((B, N, D3 // 3), dtype=torch.bfloat16, device=qkv.device)
    cls = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device)
    op = torch.empty((B, N, rollout_operand_ld(N)), dtype=torch.bfloat16, device=qkv.device)
    nbytes = int(_lib.load().vtc_attention_mean_scratch_bytes(B, N, heads))
    scratch = torch.empty((nbytes,), dtype=torch.uint8, device=qkv.device)
    _call("vtc_attention_mean_operand", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(key_bias, torch.float32, "key_bias"), _ptr(out), _ptr(cls), None,
          _ptr(op), _ptr(scratch), nbytes, B, N, heads, scale)
    return out, cls, op


def cls_layer_map(cls_rows: torch.Tensor, first: int, last: int) -> torch.Tensor:
    L, B, H, N = cls_rows.shape
    out = torch.empty((B, N - 1), dtype=torch.float32, device=cls_rows.device)
    _call("vtc_cls_layer_map", _ptr(cls_rows, torch.float32), _ptr(out), L, first, last, B, H, N)
    return out


def cam_project(tokens: torch.Tensor, w: torch.Tensor, relu: bool = True, eps: float = 1e-5) -> torch.Tensor:
    B, N, D = tokens.shape
    Ccls = w.shape[0]
    g = int(round((N - 1) ** 0.5))
    out = torch.empty((B, Ccls, g, g), dtype=torch.float32, device=tokens.device)
    _call("vtc_cam_project", _ptr(tokens, torch.float32, "tokens"), _ptr(w, torch.float32, "w"), _ptr(out), B, N, D, Ccls,
              int(relu), eps)
    return out


def normalize_max_(maps: torch.Tensor) -> torch.Tensor:
    P = maps.shape[-1]
    _call("vtc_normalize_max", _ptr(maps, torch.float32), maps.numel() // P, P)
    return maps


def upsample_bilinear(maps: torch.Tensor, out_hw: Tuple[int, int], as_u8: bool = False) -> torch.Tensor:
    """maps [..., g, g] fp32 -> [..., H, W] (fp32, or uint8 = trunc(255*v))."""
    g = maps.shape[-1]
    n = maps.numel() // (g * g)
    H, W = out_hw
    out = torch.empty((*maps.shape[:-2], H, W), dtype=torch.uint8 if as_u8 else torch.float32, device=maps.device)
    _call("vtc_upsample_bilinear_u8" if as_u8 else "vtc_upsample_bilinear", _ptr(maps, torch.float32), _ptr(out), n, g, H, W)
    return out


def cam_label(cam: torch.Tensor, labels: torch.Tensor, out_hw: Tuple[int, int], bg_thresh: float = 0.25) -> torch.Tensor:
    B, Ccls, g, _ = cam.shape
    H, W = out_hw
    lab = labels.to(torch.uint8).contiguous()
    out = torch.empty((B, H, W), dtype=torch.uint8, device=cam.device)
    _call("vtc_cam_label", _ptr(cam, torch.float32), _ptr(lab, torch.uint8), bg_thresh, _ptr(out), B, Ccls, g, H, W)
    return out


def hwp_cos_vote(hwp_logits: torch.Tensor, head1_w: torch.Tensor, hwp_tokens: torch.Tensor, tokens: torch.Tensor,
                 sig_thresh: float = 0.9) -> Tuple[torch.Tensor, torch.Tensor]:
    B, N, D = tokens.shape
    K = hwp_tokens.shape[1]
    Ccls = head1_w.shape[0]
    g = int(round((N - 1) ** 0.5))
    p2c = torch.empty((B, K), dtype=torch.int32, device=tokens.device)
    cos = torch.empty((B, K, g, g), dtype=torch.float32, device=tokens.device)
    _call("vtc_hwp_cos_vote", _ptr(hwp_logits, torch.float32), _ptr(head1_w, torch.float32), _ptr(hwp_tokens, torch.float32),
              _ptr(tokens, torch.float32), sig_thresh, _ptr(p2c), _ptr(cos), B, N, D, Ccls, K)
    return p2c, cos


def hwp_seg(cos: torch.Tensor, p2c: torch.Tensor, bg_map: torch.Tensor, out_hw: Tuple[int, int], cos_thresh: float = 0.5,
            bg_thresh: float = 0.05) -> torch.Tensor:
    B, K, g, _ = cos.shape
    H, W = out_hw
    out = torch.empty((B, H, W), dtype=torch.uint8, device=cos.device)
    _call("vtc_hwp_seg", _ptr(cos, torch.float32), _ptr(p2c, torch.int32), _ptr(bg_map, torch.float32), cos_thresh, bg_thresh,
              _ptr(out), B, K, g, H, W)
    return out


def confmat_update(mat: torch.Tensor, gt: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    n = mat.shape[0]
    _call("vtc_confmat_update", _ptr(gt, torch.uint8, "gt"), _ptr(pred, torch.uint8, "pred"), gt.numel(), n,
              _ptr(mat, torch.int64, "mat"))
    return mat


def average_precision(labels: torch.Tensor, scores: torch.Tensor, acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-image sklearn-style AP on the device: labels / scores [B,C] fp32 -> ap [B] fp64 (-1 = no positive label);
    `acc` (fp64 [2], optional) accumulates (sum of APs, number of scored images)."""
    B, C = scores.shape
    ap = torch.empty((B,), dtype=torch.float64, device=scores.device)
    _call("vtc_average_precision", _ptr(labels, torch.float32, "labels"), _ptr(scores, torch.float32, "scores"), B, C, _ptr(ap),
              _ptr(acc, torch.float64, "acc"))
    return ap


def patch_similarity(tokens: torch.Tensor) -> torch.Tensor:
    """predict.py:191-199: gram matrix of F.normalize(tokens) (default dim=1: per feature, across tokens).  [B,N,D] -> [B,N,N]."""
    B, N, D = tokens.shape
    scratch = torch.empty((B, D), dtype=torch.float32, device=tokens.device)
    sim = torch.empty((B, N, N), dtype=torch.float32, device=tokens.device)
    _call("vtc_patch_similarity", _ptr(tokens, torch.float32, "tokens"), _ptr(scratch), _ptr(sim), B, N, D)
    return sim


def patchify_u8(x: torch.Tensor, patch: int, mean, std, split: bool = False) -> torch.Tensor:
    """uint8 [B,S,S,3] -> normalised bf16 patch matrix [B*g*g, 3*patch*patch] (x 2 when split)."""
    import ctypes
    B, S, S2, Cin = x.shape
    assert S == S2 and Cin == 3 and x.dtype == torch.uint8
    g = S // patch
    out = torch.empty((B * g * g, 3 * patch * patch * (2 if split else 1)), dtype=torch.bfloat16, device=x.device)
    _call("vtc_patchify_u8", _ptr(x), (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std), _ptr(out), B, S, patch, int(split))
    return out


# ---- LayerNorm fused into the GEMMs (bf16 forward) -------------------------------------------------------------------------
def fold_ln(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """nn.Linear weight [N,K] behind a LayerNorm(gamma, beta) -> (W' bf16 [N,K], g [N], c [N]), see include/vtc.h."""
    N, K = w.shape
    wf = torch.empty((N, K), dtype=torch.bfloat16, device=w.device)
    g = torch.empty((N,), dtype=torch.float32, device=w.device)
    c = torch.empty((N,), dtype=torch.float32, device=w.device)
    _call("vtc_fold_ln", _ptr(w, torch.float32, "w"), _ptr(gamma, torch.float32), _ptr(beta, torch.float32), _ptr(bias, torch.float32),
              _ptr(wf), _ptr(g), _ptr(c), N, K)
    return wf, g, c


def residual_prep(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 [rows,D] -> (bf16 copy, stats [rows, D/128, 2])."""
    rows, D = x.shape
    xb = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    stats = torch.empty((rows, D // 128, 2), dtype=torch.float32, device=x.device)
    _call("vtc_residual_prep", _ptr(x, torch.float32, "x"), _ptr(xb), _ptr(stats), rows, D)
    return xb, stats


def gemm_resid_ln(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, residual: torch.Tensor,
                  out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """out = residual + a.w^T + bias (fp32), bf16(out), stats of out."""
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ob = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    stats = torch.empty((M, N // 128, 2), dtype=torch.float32, device=a.device)
    _call("vtc_gemm_resid_ln", _ptr(a, torch.bfloat16, "a"), _ptr(w, torch.bfloat16, "w"), _ptr(bias, torch.float32), _ptr(residual, torch.float32),
              _ptr(out, torch.float32), _ptr(ob), _ptr(stats), M, N, K)
    return out, ob, stats


def gemm_lnfold(a: torch.Tensor, wf: torch.Tensor, c: torch.Tensor, g: torch.Tensor, stats: torch.Tensor, eps: float, gelu: bool = False) -> torch.Tensor:
    """bf16 [M,N] = [GELU](LN(t).W^T + b) evaluated as rstd (a.W'^T - mean g) + c with a = bf16(t)."""
    M, K = a.shape
    N = wf.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _call("vtc_gemm_lnfold", _ptr(a, torch.bfloat16, "a"), _ptr(wf, torch.bfloat16, "wf"), _ptr(c, torch.float32), _ptr(g, torch.float32),
              _ptr(stats, torch.float32), eps, _ptr(out), M, N, K, int(gelu))
    return out
