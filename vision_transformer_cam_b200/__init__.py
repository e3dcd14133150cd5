"""B200-native ViT forward + CAM / attention-rollout path (drop-in for Jingfeng-Tang/vision_transformer_cam's
vit_model.py and the post-processing of predict.py / validate.py).  Compute = libvtc.so (hand-written sm_100a CUDA
behind the C-ABI of include/vtc.h); PyTorch only owns memory, streams and torch.distributed."""
from . import _lib  # noqa: F401
from .vit_model import (VisionTransformer, Block, Attention, Mlp, PatchEmbed, DropPath, drop_path, CamForward,  # noqa: F401
                        vit_base_patch16_224, vit_base_patch16_224_in21k, vit_base_patch32_224,
                        vit_base_patch32_224_in21k, vit_large_patch16_224, vit_large_patch16_224_in21k,
                        vit_large_patch32_224_in21k, vit_huge_patch14_224_in21k)

__version__ = "0.1.0"
