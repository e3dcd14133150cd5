"""Launch every kernel of the path at the bench shapes between cudaProfilerStart / Stop, for ONE ncu capture
(`ncu --profile-from-start off --set full ...`, see tools/README.md).  Range A = one fused forward (ViT-B/16-224, B = 256,
depth 6 in the mask-firing 'masked' regime: layers 0-4 run the plain attention kernel, layer 5 the masked one); range B = the
kernels the forward does not reach in that configuration, through their C-ABI entry points, at the same batch.

    python tools/kernel_zoo.py [batch]        (a small batch, e.g. 8, is what a compute-sanitizer pass uses)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_transformer_cam_b200 as V
from vision_transformer_cam_b200 import cam as CAM, ops
from oracle import vit_forward as VF

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = V.VisionTransformer(img_size=224, patch_size=16, embed_dim=768, depth=6, num_heads=12, num_classes=20).eval()
model.load_state_dict(VF.masked({k: v.clone() for k, v in model.state_dict().items()}, qk_scale=3.0))
model = model.to(dev)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((B, 3, 224, 224), generator=g, device=dev)
u8 = torch.randint(0, 256, (B, 224, 224, 3), generator=g, device=dev, dtype=torch.uint8)
hw = (375, 500)
labels = torch.zeros((B, 20), device=dev)
labels[torch.arange(B), torch.arange(B) % 20] = 1
qkv = torch.randn((B, 197, 3 * 768), generator=g, device=dev).bfloat16()
mean12 = torch.randn((12, B, 197, 197), generator=g, device=dev).softmax(-1)
opnd = ops.rollout_operand_from_mean(mean12)
n32 = min(B, 32)
gt = torch.randint(0, 21, (n32, 375, 500), generator=g, device=dev, dtype=torch.uint8)


def range_a():
    return model.forward_cam(x, bg=True)


def range_b(o):
    ops.attention_mean_operand(qkv, 12, 0.125)                       # attention_cs<packed P> + head_mean_packed<operand>
    ops.attention_mean(qkv, 12, 0.125)                               # ... + head_mean_packed<fp32>
    ops.rollout_operands(opnd, 197)                                  # streaming rollout
    ops.rollout(mean12[:, :min(B, 64)].contiguous())                         # fp32 entry point
    ops.patchify_u8(u8, 16, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    cam = CAM.classic_cam(o.tokens_last, model.head1.weight.data)    # cam_project
    CAM.cam_pseudo_label(cam, labels, hw)                            # cam_label (fused upsample + argmax)
    CAM.cam_upsampled(cam[:32], hw)                                  # upsample fp32
    CAM.cam_upsampled(cam[:32], hw, as_u8=True)                      # upsample u8
    CAM.layer_maps(o.cls_rows)                                       # cls_layer_map
    seg = CAM.hwp_pseudo_seg(o, model.head1.weight.data, hw)         # hwp_cos_vote, cls_layer_map, hwp_seg
    cm = CAM.ConfusionMatrix(20, device=dev)
    cm.update(gt, seg[:n32])                                          # confmat
    CAM.average_precision(labels, torch.sigmoid(o.hwp_logits))
    CAM.patch_similarity(o.tokens_last[:32])
    ops.normalize_max_(torch.rand((B, 196), device=dev))
    ops.attention(qkv[:32].contiguous(), 12, 0.125, want_attn=True)  # full-P kernel of the 6-tuple path
    ops.head_mean(torch.rand((16, 12, 197, 197), device=dev))
    ops.split_bf16(torch.rand((4096, 768), device=dev))
    torch.cuda.synchronize()


o = range_a()
range_b(o)
torch.cuda.synchronize()
torch.cuda.profiler.start()
o = range_a()
range_b(o)
torch.cuda.profiler.stop()
torch.cuda.synchronize()
print("zoo ok: bg fraction after layers 4, 5:", [round(float(o.bg[l].float().mean()), 3) for l in (4, 5)])
