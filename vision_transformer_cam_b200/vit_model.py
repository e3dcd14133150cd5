"""Drop-in replacement for the reference `vit_model.py` on B200.

Same public surface (`VisionTransformer`, `Block`, `Attention`, `Mlp`, `PatchEmbed`, `DropPath`, `drop_path`,
`_init_vit_weights`, the eight `vit_*` factories), same constructor arguments, same parameter names / shapes /
state_dict layout (incl. the unused `norm1[256]`, `norm2[32]` and the real `head1`), same RNG consumption at
construction (so `torch.manual_seed(s); create_model(...)` yields the reference's weights), and the same 6-tuple
from `forward` (reference vit_model.py:411-424).  The compute is not PyTorch: every forward runs the hand-written
sm_100a kernels of libvtc.so through the C-ABI (`include/vtc.h`); there is no eager / CPU fallback.

Differences, all deliberate: no import-time matplotlib / palette.json side effects; the
hard-coded 197 tokens / 12 heads of the reference are generalised to N / H; inference only (no autograd graph).
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from dataclasses import dataclass
from functools import partial
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops


def drop_path(x, drop_prob: float = 0., training: bool = False):
    """Stochastic depth per sample (reference vit_model.py:20-36).  Every factory uses rate 0, so this is the identity
    on the fused path; kept for API compatibility."""
    if drop_prob == 0. or not training:
        return x
    keep = 1.0 - drop_prob
    mask = torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device).add_(keep).floor_()
    return x.div(keep) * mask


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)


# ---- per-module bf16 weight cache (module-level forwards only; the fused model packs its own) ---------------------
def _bf16(param: torch.Tensor, cache: dict) -> torch.Tensor:
    key = (param.data_ptr(), param._version, param.device)
    hit = cache.get("w")
    if hit is None or hit[0] != key:
        cache["w"] = (key, ops.cast_bf16(param.detach().reshape(param.shape[0], -1).contiguous()))
    return cache["w"][1]


def _require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: input is on {x.device}; this implementation runs on sm_100a CUDA devices only "
                           "(no CPU fallback) -- move the model and the input to the GPU")


def _linear_f32(x2d_bf16: torch.Tensor, lin: nn.Linear, cache: dict, gelu: bool = False, out_f32: bool = True) -> torch.Tensor:
    w = _bf16(lin.weight, cache)
    bias = lin.bias.detach() if lin.bias is not None else torch.zeros(lin.out_features, device=w.device)
    if not out_f32:
        return ops.gemm_bf16(x2d_bf16, w, bias, _lib.EPI_BIAS_GELU if gelu else _lib.EPI_BIAS)
    zero = torch.zeros((x2d_bf16.shape[0], lin.out_features), dtype=torch.float32, device=w.device)
    return ops.gemm_bf16(x2d_bf16, w, bias, _lib.EPI_BIAS_RESIDUAL, residual=zero, out=zero)


class PatchEmbed(nn.Module):
    """Image -> patch tokens (reference vit_model.py:51-83): a k=s=patch conv, i.e. patchify + GEMM."""

    def __init__(self, img_size=224, patch_size=16, in_c=3, embed_dim=768, norm_layer=None):
        super().__init__()
        img_size = (img_size, img_size)
        patch_size = (patch_size, patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_c, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()
        self._cache: dict = {}

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        _require_cuda(x, "PatchEmbed")
        patches = ops.patchify(x.detach().float().contiguous(), self.patch_size[0])
        w = _bf16(self.proj.weight, self._cache)
        D = w.shape[0]
        zero = torch.zeros((patches.shape[0], D), dtype=torch.float32, device=x.device)
        out = ops.gemm_bf16(patches, w, self.proj.bias.detach(), _lib.EPI_BIAS_RESIDUAL, residual=zero, out=zero)
        return self.norm(out.view(B, self.num_patches, D))


class Attention(nn.Module):
    """Multi-head attention with the reference's additive background mask (vit_model.py:86-140)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop_ratio=0., proj_drop_ratio=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop_ratio)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop_ratio)
        self._cq: dict = {}
        self._cp: dict = {}

    def forward(self, x, current_layer, mask_indices):
        """x [B,N,C] fp32 -> (out [B,N,C] fp32, weights [B,H,N,N] fp32).  `mask_indices` [B,N,N] is the reference's
        additive mask -100*min(v_i+v_j,1); it is applied for current_layer > 4 exactly as a per-key bias (row 0 of the
        mask, v_0 = 0; rows with v_i = 1 only receive a softmax-invariant constant)."""
        _require_cuda(x, "Attention")
        B, N, C = x.shape
        xb = ops.cast_bf16(x.detach().float().contiguous()).view(B * N, C)
        qkv = _linear_f32(xb, self.qkv, self._cq, out_f32=False).view(B, N, 3 * C)
        kb = None
        if current_layer > 4 and mask_indices is not None:
            kb = mask_indices[:, 0, :].to(device=x.device, dtype=torch.float32).contiguous()
        o, _, weights = ops.attention(qkv, self.num_heads, float(self.scale), key_bias=kb, want_cls=False, want_attn=True)
        out = _linear_f32(o.view(B * N, C), self.proj, self._cp).view(B, N, C)
        return out, weights


class Mlp(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (reference vit_model.py:143-164); GELU is fused into the fc1 GEMM epilogue."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self._c1: dict = {}
        self._c2: dict = {}

    def forward(self, x):
        _require_cuda(x, "Mlp")
        shp = x.shape
        xb = ops.cast_bf16(x.detach().float().contiguous()).view(-1, shp[-1])
        h = _linear_f32(xb, self.fc1, self._c1, gelu=True, out_f32=False)
        return _linear_f32(h, self.fc2, self._c2).view(*shp[:-1], -1)


class Block(nn.Module):
    """Pre-norm transformer block (reference vit_model.py:167-200)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_ratio=0., attn_drop_ratio=0.,
                 drop_path_ratio=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                              attn_drop_ratio=attn_drop_ratio, proj_drop_ratio=drop_ratio)
        self.drop_path = DropPath(drop_path_ratio) if drop_path_ratio > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop_ratio)

    def forward(self, x, current_layer, mask_indices):
        _require_cuda(x, "Block")
        B, N, C = x.shape
        x = x.detach().float().contiguous()
        a = self.attn
        y = ops.layernorm_bf16(x, self.norm1.weight.detach(), self.norm1.bias.detach(), self.norm1.eps).view(B * N, C)
        qkv = _linear_f32(y, a.qkv, a._cq, out_f32=False).view(B, N, 3 * C)
        kb = None
        if current_layer > 4 and mask_indices is not None:
            kb = mask_indices[:, 0, :].to(device=x.device, dtype=torch.float32).contiguous()
        o, _, weights = ops.attention(qkv, a.num_heads, float(a.scale), key_bias=kb, want_cls=False, want_attn=True)
        x1 = ops.gemm_bf16(o.view(B * N, C), _bf16(a.proj.weight, a._cp), a.proj.bias.detach(), _lib.EPI_BIAS_RESIDUAL,
                           residual=x.view(B * N, C))
        y2 = ops.layernorm_bf16(x1, self.norm2.weight.detach(), self.norm2.bias.detach(), self.norm2.eps)
        h = _linear_f32(y2, self.mlp.fc1, self.mlp._c1, gelu=True, out_f32=False)
        x2 = ops.gemm_bf16(h, _bf16(self.mlp.fc2.weight, self.mlp._c2), self.mlp.fc2.bias.detach(), _lib.EPI_BIAS_RESIDUAL,
                           residual=x1, out=x1)
        return x2.view(B, N, C), weights


@dataclass
class CamForward:
    """Compact outputs of `VisionTransformer.forward_cam` (everything the CAM / rollout / pseudo-label code consumes)."""
    logits: torch.Tensor                 # [B,C]
    hwp_logits: torch.Tensor             # [B,C]
    hwp_tokens: torch.Tensor             # [B,K,D]
    topk_idx: torch.Tensor               # [B,K] int32
    tokens: torch.Tensor                 # [Lt,B,N,D] trailing block outputs (tokens[-1] = block-L output)
    cls_rows: torch.Tensor               # [L,B,H,N]
    attn: Optional[torch.Tensor] = None       # [La,B,H,N,N]
    attn_mean: Optional[torch.Tensor] = None  # [L,B,N,N]
    bg: Optional[torch.Tensor] = None         # [L,B,P] uint8
    cls_map: Optional[torch.Tensor] = None    # [L,B,P]
    rollout: Optional[torch.Tensor] = None    # [B,P] un-normalised attention-rollout row over the last min(L,12) layers

    @property
    def tokens_last(self) -> torch.Tensor:
        return self.tokens[-1]


class _Engine:
    """Owns the libvtc model handle, the packed bf16 weights and the workspace for one VisionTransformer."""

    def __init__(self, model: "VisionTransformer"):
        self.lib = _lib.load()
        pe = model.patch_embed
        cfg = _lib.Config(img_size=pe.img_size[0], patch_size=pe.patch_size[0], in_c=pe.proj.in_channels,
                          num_classes=model.num_classes, embed_dim=model.embed_dim, depth=len(model.blocks),
                          num_heads=model.blocks[0].attn.num_heads, mlp_hidden=model.blocks[0].mlp.fc1.out_features,
                          representation_size=model.num_features if model.has_logits else 0,
                          mask_from=4, mask_thresh=0.25, topk=16, ln_eps=float(model.norm.eps))
        self.cfg = cfg
        handle = ctypes.c_void_p()
        _lib.check(self.lib.vtc_model_create(ctypes.byref(cfg), ctypes.byref(handle)), "vtc_model_create")
        self.handle = handle
        self.packed: Optional[torch.Tensor] = None
        self.precision = "bf16"
        self.key = None
        self.ws: Optional[torch.Tensor] = None
        self._keep = None
        self._params = None
        self.graphs: Dict[tuple, tuple] = {}          # forward_cam_graphed: key -> (CUDAGraph, static input, outputs, workspace, pointer key)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.vtc_model_destroy(self.handle)
        except Exception:
            pass

    def profile(self, enable: bool) -> None:
        _lib.check(self.lib.vtc_model_profile(self.handle, int(enable)), "vtc_model_profile")

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """Per-kernel-kind (milliseconds, launches) accumulated since the last read (synchronises on the recorded events)."""
        n = len(_lib.PROF_KINDS)
        ms = (ctypes.c_float * n)()
        cnt = (ctypes.c_int32 * n)()
        _lib.check(self.lib.vtc_model_profile_read(self.handle, ms, cnt), "vtc_model_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.PROF_KINDS)}

    def ensure_packed(self, model: "VisionTransformer", device: torch.device) -> None:
        if self._params is None:
            self._params = dict(model.named_parameters())
        params = self._params
        key = (device, model.precision, tuple((p.data_ptr(), p._version) for p in params.values()))
        if key == self.key:
            return
        self._params = params = dict(model.named_parameters())       # slow path: the module may have been restructured
        key = (device, model.precision, tuple((p.data_ptr(), p._version) for p in params.values()))
        if model.precision != self.precision:
            _lib.check(self.lib.vtc_model_set_precision(self.handle, _lib.PRECISION_FP32_SPLIT if model.precision == "fp32"
                                                        else _lib.PRECISION_BF16), "vtc_model_set_precision")
            self.precision = model.precision
        for n, p in params.items():
            if p.device != device:
                raise RuntimeError(f"parameter {n} is on {p.device}, input on {device}")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"parameter {n} must be contiguous fp32")
        L = len(model.blocks)
        layers = (_lib.LayerWeights * L)()
        for i, blk in enumerate(model.blocks):
            lw = layers[i]
            lw.norm1_w, lw.norm1_b = blk.norm1.weight.data_ptr(), blk.norm1.bias.data_ptr()
            lw.qkv_w, lw.qkv_b = blk.attn.qkv.weight.data_ptr(), blk.attn.qkv.bias.data_ptr()
            lw.proj_w, lw.proj_b = blk.attn.proj.weight.data_ptr(), blk.attn.proj.bias.data_ptr()
            lw.norm2_w, lw.norm2_b = blk.norm2.weight.data_ptr(), blk.norm2.bias.data_ptr()
            lw.fc1_w, lw.fc1_b = blk.mlp.fc1.weight.data_ptr(), blk.mlp.fc1.bias.data_ptr()
            lw.fc2_w, lw.fc2_b = blk.mlp.fc2.weight.data_ptr(), blk.mlp.fc2.bias.data_ptr()
        w = _lib.Weights()
        w.cls_token, w.pos_embed = model.cls_token.data_ptr(), model.pos_embed.data_ptr()
        w.patch_w, w.patch_b = model.patch_embed.proj.weight.data_ptr(), model.patch_embed.proj.bias.data_ptr()
        w.norm_w, w.norm_b = model.norm.weight.data_ptr(), model.norm.bias.data_ptr()
        if model.has_logits:
            w.pre_w, w.pre_b = model.pre_logits.fc.weight.data_ptr(), model.pre_logits.fc.bias.data_ptr()
        w.head_w, w.head_b = model.head.weight.data_ptr(), model.head.bias.data_ptr()
        w.head1_w, w.head1_b = model.head1.weight.data_ptr(), model.head1.bias.data_ptr()
        w.layers = layers
        w.num_layers = L
        nbytes = self.lib.vtc_model_packed_bytes(self.handle)
        if self.packed is None or self.packed.device != device or self.packed.numel() < nbytes:
            self.graphs.clear()          # captured graphs point into the old buffer
            self.packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(self.lib.vtc_model_pack_weights(self.handle, ctypes.byref(w), self.packed.data_ptr(), nbytes,
                                                   torch.cuda.current_stream(device).cuda_stream), "vtc_model_pack_weights")
        self.key = key
        self._keep = (layers, w)

    def run(self, model: "VisionTransformer", x: torch.Tensor, tokens_layers: int, attn_layers: int, attn_mean: bool, bg: bool,
            cls_map: bool, mask_norm: str, forced_bg: Optional[Dict[int, torch.Tensor]], forced_topk: Optional[torch.Tensor],
            norm: Optional[Tuple[Tuple[float, float, float], Tuple[float, float, float]]] = None, rollout: bool = False) -> CamForward:
        dev = x.device
        cfg = self.cfg
        B = x.shape[0]
        g = cfg.img_size // cfg.patch_size
        P, N, D, H, L, Ccls, K = g * g, g * g + 1, cfg.embed_dim, cfg.num_heads, cfg.depth, cfg.num_classes, cfg.topk
        with torch.cuda.device(dev):
            self.ensure_packed(model, dev)
            f32 = dict(dtype=torch.float32, device=dev)
            out = CamForward(
                logits=torch.empty((B, Ccls), **f32), hwp_logits=torch.empty((B, Ccls), **f32),
                hwp_tokens=torch.empty((B, K, D), **f32), topk_idx=torch.empty((B, K), dtype=torch.int32, device=dev),
                tokens=torch.empty((tokens_layers, B, N, D), **f32), cls_rows=torch.empty((L, B, H, N), **f32))
            if attn_layers > 0:
                out.attn = torch.empty((attn_layers, B, H, N, N), **f32)
            if attn_mean:
                out.attn_mean = torch.empty((L, B, N, N), **f32)
            if bg:
                out.bg = torch.empty((L, B, P), dtype=torch.uint8, device=dev)
            if cls_map:
                out.cls_map = torch.empty((L, B, P), **f32)
            if rollout:
                out.rollout = torch.empty((B, P), **f32)
            o = _lib.Outputs(logits=out.logits.data_ptr(), hwp_logits=out.hwp_logits.data_ptr(), hwp_tokens=out.hwp_tokens.data_ptr(),
                             topk_idx=out.topk_idx.data_ptr(), tokens=out.tokens.data_ptr(), tokens_layers=tokens_layers,
                             cls_rows=out.cls_rows.data_ptr(), attn=out.attn.data_ptr() if out.attn is not None else None,
                             attn_layers=attn_layers, attn_mean=out.attn_mean.data_ptr() if attn_mean else None,
                             bg=out.bg.data_ptr() if bg else None, cls_map=out.cls_map.data_ptr() if cls_map else None,
                             rollout=out.rollout.data_ptr() if rollout else None)
            forcing = None
            keep = []
            if forced_bg or forced_topk is not None:
                forcing = _lib.Forcing()
                if forced_bg:
                    buf = torch.zeros((L, B, P), dtype=torch.uint8, device=dev)
                    mask = 0
                    for l, v in forced_bg.items():
                        buf[l] = v.to(device=dev, dtype=torch.uint8)
                        mask |= 1 << l
                    forcing.bg, forcing.bg_layer_mask = buf.data_ptr(), mask
                    keep.append(buf)
                if forced_topk is not None:
                    tk = forced_topk.to(device=dev, dtype=torch.int32).contiguous()
                    forcing.topk_idx = tk.data_ptr()
                    keep.append(tk)
            need = self.lib.vtc_workspace_bytes(self.handle, B, ctypes.byref(o))
            if self.ws is None or self.ws.device != dev or self.ws.numel() < need:
                self.ws = None
                self.ws = torch.empty(need, dtype=torch.uint8, device=dev)
            flags = _lib.FWD_MASK_NORM_IMAGE if mask_norm == "image" else 0
            if self.precision == "fp32":
                flags |= _lib.FWD_FP32_SPLIT
            fptr = ctypes.byref(forcing) if forcing is not None else None
            stream = torch.cuda.current_stream(dev).cuda_stream
            if norm is None:
                _lib.check(self.lib.vtc_forward(self.handle, x.data_ptr(), B, ctypes.byref(o), fptr, self.ws.data_ptr(),
                                                self.ws.numel(), flags, stream), "vtc_forward")
            else:
                mean, std = (ctypes.c_float * 3)(*norm[0]), (ctypes.c_float * 3)(*norm[1])
                _lib.check(self.lib.vtc_forward_u8(self.handle, x.data_ptr(), mean, std, B, ctypes.byref(o), fptr, self.ws.data_ptr(),
                                                   self.ws.numel(), flags, stream), "vtc_forward_u8")
            del keep
        return out


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_c=3, num_classes=1000,
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, representation_size=None, distilled=False, drop_ratio=0.,
                 attn_drop_ratio=0., drop_path_ratio=0., embed_layer=PatchEmbed, norm_layer=None,
                 act_layer=None, is_train=True):
        """Arguments as in the reference (vit_model.py:215-219).  `distilled` models are not supported (the reference's
        own forward cannot run them); dropout / drop-path rates must be 0 (every factory's value)."""
        super().__init__()
        if distilled:
            raise NotImplementedError("distilled ViT: unreachable in the reference forward (vit_model.py:404-412)")
        if drop_ratio or attn_drop_ratio or drop_path_ratio:
            raise NotImplementedError("non-zero dropout / drop-path: the fused inference path has no stochastic ops")
        if not qkv_bias:
            raise NotImplementedError("qkv_bias=False is not used by any reference factory")
        # the fused forward computes scale = head_dim ** -0.5, erf-GELU and LayerNorm; anything else would be silently ignored
        if qk_scale is not None and abs(float(qk_scale) - (embed_dim // num_heads) ** -0.5) > 1e-12:
            raise NotImplementedError("qk_scale other than head_dim ** -0.5 is not used by any reference factory")
        if act_layer is not None and act_layer is not nn.GELU:
            raise NotImplementedError("act_layer other than nn.GELU (erf) is not supported by the fused forward")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 1
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU

        # construction order == the reference's, so the RNG stream (default nn inits) is consumed identically
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_c=in_c, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = None
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + self.num_tokens, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_ratio)
        self.blocks = nn.Sequential(*[
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop_ratio=drop_ratio, attn_drop_ratio=attn_drop_ratio, drop_path_ratio=0.,
                  norm_layer=norm_layer, act_layer=act_layer)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        if not (type(self.norm) is nn.LayerNorm and self.norm.elementwise_affine and self.norm.bias is not None):
            raise NotImplementedError("norm_layer must build an affine nn.LayerNorm (any eps): the fused forward has no other norm")
        if representation_size:
            self.has_logits = True
            self.num_features = representation_size
            self.pre_logits = nn.Sequential(OrderedDict([("fc", nn.Linear(embed_dim, representation_size)), ("act", nn.Tanh())]))
        else:
            self.has_logits = False
            self.pre_logits = nn.Identity()
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.head_dist = None

        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        self.apply(_init_vit_weights)

        # created after the init pass -> default nn init, exactly like the reference (vit_model.py:292-295)
        self.norm1 = norm_layer(256)
        self.norm2 = norm_layer(32)
        self.head1 = nn.Linear(self.num_features, num_classes)
        self.relu = nn.ReLU()
        self.is_train = is_train
        self.final_seg_count = 0
        self.precision = "bf16"
        self._engine: Optional[_Engine] = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None          # the libvtc handle is process-local; it is rebuilt lazily
        return state

    def set_precision(self, precision: str) -> "VisionTransformer":
        """'bf16' (default): bf16 tensor-core operands, logits within 1e-2 of the fp32 reference.  'fp32': every GEMM /
        attention operand is carried as a (hi, lo) bf16 pair and every product evaluated as hi.hi + lo.hi + hi.lo on the
        tensor cores (3x the MMA work), logits within 1e-4 (BASELINE.json north_star "fp32 mode")."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        return self

    # -- fused path -----------------------------------------------------------------------------------------------
    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        B, C, H, W = x.shape
        isz = self.patch_embed.img_size
        assert H == isz[0] and W == isz[1], f"Input image size ({H}*{W}) doesn't match model ({isz[0]}*{isz[1]})."
        _require_cuda(x, "VisionTransformer")
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        return x.contiguous()

    @torch.no_grad()
    def forward_cam(self, x: torch.Tensor, tokens_layers: int = 1, attn_layers: int = 0, attn_mean: bool = False,
                    bg: bool = False, cls_map: bool = False, mask_norm: str = "batch",
                    forced_bg: Optional[Dict[int, torch.Tensor]] = None, forced_topk: Optional[torch.Tensor] = None,
                    rollout: bool = False) -> CamForward:
        """Fused forward with compact outputs (no [B,H,N,N] tensors unless `attn_layers` > 0).  `rollout=True` also returns the
        attention-rollout row of predict.py:215-232 ([B,P], un-normalised) computed inside the call: the head means never
        leave the workspace (bf16 rollout operands, half the bytes of `attn_mean`)."""
        assert mask_norm in ("batch", "image")
        x = self._check_input(x)
        if self._engine is None:
            self._engine = _Engine(self)
        return self._engine.run(self, x, tokens_layers, attn_layers, attn_mean, bg, cls_map, mask_norm, forced_bg, forced_topk, rollout=rollout)

    @torch.no_grad()
    def forward_cam_u8(self, x: torch.Tensor, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), **kw) -> CamForward:
        """`forward_cam` fed with decoded images: x uint8 [B,S,S,3] (HWC, what PIL / cv2 decode to) on the GPU.  ToTensor
        (/255) and Normalize(mean, std) of the reference's transforms (predict.py:72-75, validate.py:80-84) run inside the
        patch-matrix kernel with the same fp32 operations, so the result equals forward_cam(Normalize(ToTensor(x))) bit for
        bit while moving 4x fewer input bytes."""
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError(f"forward_cam_u8 expects uint8 [B,S,S,3], got {x.dtype} {tuple(x.shape)}")
        isz = self.patch_embed.img_size
        assert x.shape[1] == isz[0] and x.shape[2] == isz[1], f"Input image size ({x.shape[1]}*{x.shape[2]}) doesn't match model ({isz[0]}*{isz[1]})."
        _require_cuda(x, "VisionTransformer")
        kw.setdefault("tokens_layers", 1)
        if self._engine is None:
            self._engine = _Engine(self)
        return self._engine.run(self, x.detach().contiguous(), kw.pop("tokens_layers"), kw.pop("attn_layers", 0), kw.pop("attn_mean", False),
                                kw.pop("bg", False), kw.pop("cls_map", False), kw.pop("mask_norm", "batch"), kw.pop("forced_bg", None),
                                kw.pop("forced_topk", None), norm=(tuple(mean), tuple(std)), rollout=kw.pop("rollout", False))

    @torch.no_grad()
    def forward_cam_graphed(self, x: torch.Tensor, **kw) -> CamForward:
        """`forward_cam` replayed from a CUDA graph: for the launch-bound small batches of the reference's own drivers
        (predict.py and validate.py run batch 1: ~105 kernel launches for ~0.3 ms of GPU work).  The first call for a given
        (shape, dtype, options) runs the forward twice eagerly and captures it; later calls copy `x` into the graph's input
        buffer and replay.  The returned tensors are the graph's static outputs: they are overwritten by the next call with
        the same signature (clone what must survive).  Weight updates are picked up (packed copies are refreshed before the
        replay); a change of parameter storage re-captures.  Options: those of forward_cam except the forced_* debugging aids."""
        if "forced_bg" in kw or "forced_topk" in kw:
            raise ValueError("forward_cam_graphed does not take forced_bg / forced_topk")
        u8 = x.dtype == torch.uint8
        if not u8:
            x = self._check_input(x)
        else:
            _require_cuda(x, "VisionTransformer")
        if self._engine is None:
            self._engine = _Engine(self)
        eng = self._engine
        dev = x.device
        call = self.forward_cam_u8 if u8 else self.forward_cam
        gkey = (tuple(x.shape), x.dtype, dev, self.precision, tuple(sorted(kw.items())))
        with torch.cuda.device(dev):
            eng.ensure_packed(self, dev)
            ptrs = (eng.packed.data_ptr(),) + tuple(k[0] for k in eng.key[2])
            ent = eng.graphs.get(gkey)
            if ent is not None and ent[4] != ptrs:
                ent = None                                   # parameters or the packed weights moved: the captured pointers are stale
            if ent is None:
                static_x = x.clone()
                cur = torch.cuda.current_stream(dev)
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):                # warm-up off the capture: function attributes, workspace growth
                    for _ in range(2):
                        call(static_x, **kw)
                cur.wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = call(static_x, **kw)
                ent = (graph, static_x, out, eng.ws, ptrs, eng.packed)   # workspace and packed weights the graph points into stay alive with the entry
                eng.graphs[gkey] = ent
            graph, static_x, out = ent[0], ent[1], ent[2]
            static_x.copy_(x, non_blocking=True)
            graph.replay()
        return out

    @torch.no_grad()
    def topk_heads(self, tokens: torch.Tensor, cls_map: torch.Tensor, gmax: Optional[torch.Tensor] = None,
                   forced_topk: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """vit_model.py:372-422 on given inputs (`vtc_topk_heads`): tokens [B,N,D] fp32 (block-L output), cls_map [B,P] (the
        renormalised CLS row of :366-371), gmax [1] (batch-global max, :372; None = per-image max)
        -> (logits [B,C], hwp_logits [B,C], hwp_tokens [B,16,D], topk_idx [B,16] int32)."""
        _require_cuda(tokens, "topk_heads")
        if self._engine is None:
            self._engine = _Engine(self)
        eng, dev = self._engine, tokens.device
        B, K, D, C = tokens.shape[0], eng.cfg.topk, eng.cfg.embed_dim, eng.cfg.num_classes
        with torch.cuda.device(dev):
            eng.ensure_packed(self, dev)
            f32 = dict(dtype=torch.float32, device=dev)
            logits, hwp, ori = torch.empty((B, C), **f32), torch.empty((B, C), **f32), torch.empty((B, K, D), **f32)
            idx = torch.empty((B, K), dtype=torch.int32, device=dev)
            tk = None if forced_topk is None else forced_topk.to(device=dev, dtype=torch.int32).contiguous()
            tokens, cls_map = tokens.float().contiguous(), cls_map.float().contiguous()
            _lib.check(eng.lib.vtc_topk_heads(eng.handle, tokens.data_ptr(), cls_map.data_ptr(), None if gmax is None else gmax.data_ptr(),
                                              None if tk is None else tk.data_ptr(), logits.data_ptr(), hwp.data_ptr(), ori.data_ptr(),
                                              idx.data_ptr(), B, torch.cuda.current_stream(dev).cuda_stream), "vtc_topk_heads")
        return logits, hwp, ori, idx

    def kernel_profile(self, enable: Optional[bool] = None):
        """enable/disable CUDA-event timing of every kernel of the fused forward, or (no argument) read the accumulated
        {kind: (ms, launches)} table.  Used by bench.py for the roofline numbers."""
        if self._engine is None:
            self._engine = _Engine(self)
        if enable is None:
            return self._engine.profile_read()
        self._engine.profile(enable)
        return None

    def forward(self, x):
        """Reference-compatible 6-tuple (vit_model.py:424): (logits [B,C], attn_weights list of [B,H,N,N], attn_matrix list
        of [B,N,D], hwp logits [B,C], head1.weight.data [C,D], hwp tokens [B,16,D]); the lists hold the last
        min(depth,12) layers (vit_model.py:322)."""
        keep = min(len(self.blocks), 12)
        o = self.forward_cam(x, tokens_layers=keep, attn_layers=keep)
        attn_weights = [o.attn[i] for i in range(keep)]
        attn_matrix = [o.tokens[i] for i in range(keep)]
        return o.logits, attn_weights, attn_matrix, o.hwp_logits, self.head1.weight.data, o.hwp_tokens


def _init_vit_weights(m):
    """ViT weight initialisation (reference vit_model.py:427-442)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.01)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_out")
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.LayerNorm):
        nn.init.zeros_(m.bias)
        nn.init.ones_(m.weight)


# ---- factories (reference vit_model.py:445-577): same names, arguments and hyper-parameters -------------------------
def _vit(patch, dim, depth, heads, num_classes, rep):
    return VisionTransformer(img_size=224, patch_size=patch, embed_dim=dim, depth=depth, num_heads=heads,
                             representation_size=rep, num_classes=num_classes)


def vit_base_patch16_224(num_classes: int = 1000):
    return _vit(16, 768, 12, 12, num_classes, None)


def vit_base_patch16_224_in21k(num_classes: int = 21843, has_logits: bool = True):
    return _vit(16, 768, 12, 12, num_classes, 768 if has_logits else None)


def vit_base_patch32_224(num_classes: int = 1000):
    return _vit(32, 768, 12, 12, num_classes, None)


def vit_base_patch32_224_in21k(num_classes: int = 21843, has_logits: bool = True):
    return _vit(32, 768, 12, 12, num_classes, 768 if has_logits else None)


def vit_large_patch16_224(num_classes: int = 1000):
    return _vit(16, 1024, 24, 16, num_classes, None)


def vit_large_patch16_224_in21k(num_classes: int = 21843, has_logits: bool = True):
    return _vit(16, 1024, 24, 16, num_classes, 1024 if has_logits else None)


def vit_large_patch32_224_in21k(num_classes: int = 21843, has_logits: bool = True):
    return _vit(32, 1024, 24, 16, num_classes, 1024 if has_logits else None)


def vit_huge_patch14_224_in21k(num_classes: int = 21843, has_logits: bool = True):
    """ViT-H/14 (257 tokens, head_dim 80): patch matrix with K padded 588 -> 640, attention through the general-shape kernel
    (csrc/attention_generic.cu); bf16 mode, fp32 image input."""
    return _vit(14, 1280, 32, 16, num_classes, 1280 if has_logits else None)
