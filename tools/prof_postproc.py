"""CUDA-event timing of the post-processing kernels at the bench batch (B = 256, VOC-shaped 375 x 500 outputs) with their
algorithmic bytes and the resulting GB/s.    python tools/prof_postproc.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import cam as CAM, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False).to(dev).eval()
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((B, 3, 224, 224), generator=g, device=dev)
u8 = torch.randint(0, 256, (B, 224, 224, 3), generator=g, device=dev, dtype=torch.uint8)
o = model.forward_cam(x, rollout=True)
hw = (375, 500)
H, W = hw
labels = torch.zeros((B, 20), device=dev)
labels[torch.arange(B), torch.arange(B) % 20] = 1
labels[torch.arange(0, B, 2), (torch.arange(0, B, 2) + 7) % 20] = 1          # 1.5 labels per image
cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)
p2c, cos = ops.hwp_cos_vote(o.hwp_logits, model.head1.weight.data, o.hwp_tokens, o.tokens_last.contiguous(), 0.9)
bgm = CAM.bg_map(o.cls_rows)
seg = ops.hwp_seg(cos, p2c, bgm, hw)
gt = torch.randint(0, 21, (B, H, W), generator=g, device=dev, dtype=torch.uint8)
mean12 = torch.randn((12, B, 197, 197), generator=g, device=dev).softmax(-1)
opnd = ops.rollout_operand_from_mean(mean12)
cm = CAM.ConfusionMatrix(20, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

cases = [
    ("patchify fp32", lambda: ops.patchify(x, 16), B * (602112 + 301056)),
    ("patchify u8", lambda: ops.patchify_u8(u8, 16, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), B * (150528 + 301056)),
    ("topk_heads", lambda: model.topk_heads(o.tokens_last, o.cls_rows[-1].mean(1)[:, 1:].contiguous(), None), B * (16 * 768 * 4 * 2 + 768 * 4)),
    ("cam_project", lambda: CAM.classic_cam(o.tokens_last, model.head1.weight.data), B * (196 * 768 * 4 + 20 * 196 * 4)),
    ("cam_label 375x500", lambda: CAM.cam_pseudo_label(cam, labels, hw), B * (20 * 196 * 4 + H * W)),
    ("upsample fp32 375x500 (32 img x 20 maps)", lambda: CAM.cam_upsampled(cam[:32], hw), 32 * 20 * (196 * 4 + H * W * 4)),
    ("upsample u8 375x500 (12 layers x B maps)", lambda: CAM.layer_maps(o.cls_rows, hw, as_u8=True), 12 * B * (196 * 4 + H * W)),
    ("hwp_cos_vote", lambda: ops.hwp_cos_vote(o.hwp_logits, model.head1.weight.data, o.hwp_tokens, o.tokens_last, 0.9), B * (197 * 768 * 4 + 16 * 768 * 4 + 16 * 196 * 4)),
    ("hwp_seg 375x500", lambda: ops.hwp_seg(cos, p2c, bgm, hw), B * (17 * 196 * 4 + H * W)),
    ("confmat 375x500", lambda: cm.update(gt, seg), B * 2 * H * W),
    ("cls_layer_map (bg map)", lambda: CAM.bg_map(o.cls_rows), B * (7 * 12 * 197 * 4 + 196 * 4)),
    ("rollout (bf16 operands, streaming)", lambda: ops.rollout_operands(opnd, 197), B * (11 * 197 * 200 * 2 + 400)),
    ("rollout (fp32 entry point)", lambda: ops.rollout(mean12), B * 12 * 197 * 197 * 4),
    ("average_precision", lambda: CAM.average_precision(labels, torch.sigmoid(o.hwp_logits)), B * 160),
    ("patch_similarity (32 img)", lambda: CAM.patch_similarity(o.tokens_last[:32]), 32 * (197 * 768 * 4 + 197 * 197 * 4)),
]
print(f"B = {B}; time = mean of 5 launches, each after an L2 flush (256 MB write); bytes = algorithmic (SURVEY D.2)")
for name, fn, nbytes in cases:
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / 5 * 1e3
    print(f"{name:44s} {us:9.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / us / 1e3:8.0f} GB/s  ({nbytes / us / 1e3 / 6545.9:.2f} of the measured HBM peak)", flush=True)
