"""ctypes binding of libvtc.so (include/vtc.h).  No torch types cross this boundary: only raw device
pointers, sizes and a cudaStream_t.  There is no fallback: a missing library or a non-sm_100 device raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VTC_LIB_PATH") or os.path.join(HERE, "libvtc.so")      # VTC_LIB_PATH: an instrumented build (tools/build_ablate.sh)

VTC_OK = 0
ERR_NAMES = {-1: "VTC_ERR_ARG", -2: "VTC_ERR_SHAPE", -3: "VTC_ERR_ARCH", -4: "VTC_ERR_CUDA", -5: "VTC_ERR_WORKSPACE"}

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_PATCH_EMBED = 0, 1, 2, 3
FWD_MASK_NORM_IMAGE = 1 << 0
FWD_FP32_SPLIT = 1 << 1
PRECISION_BF16, PRECISION_FP32_SPLIT = 0, 1
PROF_KINDS = ("patchify", "gemm_patch", "layernorm", "gemm_qkv", "attention", "gemm_proj", "gemm_fc1", "gemm_fc2", "cls", "head_mean", "heads", "rollout")

c_f32p = C.c_void_p   # device pointers are passed as integers


class VtcError(RuntimeError):
    def __init__(self, code: int, where: str, msg: str):
        super().__init__(f"{where} failed: {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("img_size", C.c_int32), ("patch_size", C.c_int32), ("in_c", C.c_int32), ("num_classes", C.c_int32),
                ("embed_dim", C.c_int32), ("depth", C.c_int32), ("num_heads", C.c_int32), ("mlp_hidden", C.c_int32),
                ("representation_size", C.c_int32), ("mask_from", C.c_int32), ("mask_thresh", C.c_float),
                ("topk", C.c_int32), ("ln_eps", C.c_float)]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("norm1_w", "norm1_b", "qkv_w", "qkv_b", "proj_w", "proj_b",
                                          "norm2_w", "norm2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class Weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cls_token", "pos_embed", "patch_w", "patch_b", "norm_w", "norm_b",
                                          "pre_w", "pre_b", "head_w", "head_b", "head1_w", "head1_b")] + \
               [("layers", C.POINTER(LayerWeights)), ("num_layers", C.c_int32)]


class Forcing(C.Structure):
    _fields_ = [("bg", C.c_void_p), ("bg_layer_mask", C.c_uint32), ("topk_idx", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("hwp_logits", C.c_void_p), ("hwp_tokens", C.c_void_p), ("topk_idx", C.c_void_p),
                ("tokens", C.c_void_p), ("tokens_layers", C.c_int32), ("cls_rows", C.c_void_p), ("attn", C.c_void_p),
                ("attn_layers", C.c_int32), ("attn_mean", C.c_void_p), ("bg", C.c_void_p), ("cls_map", C.c_void_p), ("rollout", C.c_void_p)]


_P, _I, _F, _Z, _U = C.c_void_p, C.c_int32, C.c_float, C.c_size_t, C.c_uint32

# name -> (restype, argtypes); mirrors include/vtc.h one to one (tests check the export list against the header)
SIGNATURES = {
    "vtc_version": (C.c_int, []),
    "vtc_last_error": (C.c_char_p, []),
    "vtc_check_device": (C.c_int, []),
    "vtc_launch_count": (C.c_uint64, []),
    "vtc_model_profile": (C.c_int, [_P, _I]),
    "vtc_model_profile_read": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "vtc_model_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "vtc_model_destroy": (C.c_int, [_P]),
    "vtc_model_packed_bytes": (_Z, [_P]),
    "vtc_model_pack_weights": (C.c_int, [_P, C.POINTER(Weights), _P, _Z, _P]),
    "vtc_workspace_bytes": (_Z, [_P, _I, C.POINTER(Outputs)]),
    "vtc_forward": (C.c_int, [_P, _P, _I, C.POINTER(Outputs), C.POINTER(Forcing), _P, _Z, _U, _P]),
    "vtc_forward_u8": (C.c_int, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, C.POINTER(Outputs), C.POINTER(Forcing), _P, _Z, _U, _P]),
    "vtc_gemm_bf16": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vtc_model_set_precision": (C.c_int, [_P, _I]),
    "vtc_gemm_split": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vtc_split_bf16": (C.c_int, [_P, _P, _Z, _Z, _P]),
    "vtc_patchify_split": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "vtc_layernorm_split": (C.c_int, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "vtc_patchify_u8": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _I, _I, _I, _I, _P]),
    "vtc_average_precision": (C.c_int, [_P, _P, _I, _I, _P, _P, _P]),
    "vtc_patch_similarity": (C.c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "vtc_gemm_resid_ln": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "vtc_gemm_lnfold": (C.c_int, [_P, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _P]),
    "vtc_residual_prep": (C.c_int, [_P, _P, _P, _I, _I, _P]),
    "vtc_fold_ln": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "vtc_cast_bf16": (C.c_int, [_P, _P, _Z, _P]),
    "vtc_patchify": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "vtc_cls_token_rows": (C.c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "vtc_layernorm_bf16": (C.c_int, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "vtc_attention": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "vtc_attention_mean_scratch_bytes": (_Z, [_I, _I, _I]),
    "vtc_attention_mean": (C.c_int, [_P, _P, _P, _P, _P, _P, _Z, _I, _I, _I, _F, _P]),
    "vtc_attention_generic": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "vtc_attention_kv": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P]),
    "vtc_head_mean": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    "vtc_cls_stat": (C.c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "vtc_cls_mask": (C.c_int, [_P, _P, _P, _F, _I, _P, _P, _I, _I, _P]),
    "vtc_topk_heads": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "vtc_attention_mask_operand_bytes": (_Z, [_I]),
    "vtc_cls_stat_mask": (C.c_int, [_P, _P, _P, _P, _F, _I, _P, _P, _P, _P, _F, _I, _I, _I, _P]),
    "vtc_attention_masked": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "vtc_rollout": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    "vtc_rollout_operand_ld": (C.c_int32, [_I]),
    "vtc_rollout_operand_from_mean": (C.c_int, [_P, _P, _I, _I, _P]),
    "vtc_rollout_operands": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    "vtc_attention_mean_operand": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _Z, _I, _I, _I, _F, _P]),
    "vtc_cls_layer_map": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "vtc_cam_project": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "vtc_normalize_max": (C.c_int, [_P, _I, _I, _P]),
    "vtc_upsample_bilinear": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "vtc_upsample_bilinear_u8": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "vtc_cam_label": (C.c_int, [_P, _P, _F, _P, _I, _I, _I, _I, _I, _P]),
    "vtc_hwp_cos_vote": (C.c_int, [_P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vtc_hwp_seg": (C.c_int, [_P, _P, _P, _F, _F, _P, _I, _I, _I, _I, _I, _P]),
    "vtc_confmat_update": (C.c_int, [_P, _P, _Z, _I, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libvtc.so (built in-tree by `python -m vision_transformer_cam_b200.build`).  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m vision_transformer_cam_b200.build` "
                           "(there is no CPU / eager fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().vtc_last_error().decode("utf-8", "replace")


def check(code: int, where: str) -> None:
    if code != VTC_OK:
        raise VtcError(code, where, last_error())


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
