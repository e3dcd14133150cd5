"""Run the attention kernel alone at the bench shape (B=256, N=197, H=12) for ncu / timing.
   python tools/prof_attn.py [N] [H] [B] [kv]"""
import sys, torch
sys.path.insert(0, ".")
from vision_transformer_cam_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 197
H = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
kv = len(sys.argv) > 4 and sys.argv[4] == "kv"
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
qkv = (torch.randn((B, N, 3 * H * 64), generator=g, device=dev) * 1.0).bfloat16()
kb = torch.where(torch.rand((B, N), generator=g, device=dev) < 0.3, -100.0, 0.0)
kb[:, 0] = 0
fn = ops.attention_kv if kv else ops.attention
for bias in (None, kb):
    for _ in range(3):
        fn(qkv, H, 0.125, key_bias=bias, want_cls=True, want_attn=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn(qkv, H, 0.125, key_bias=bias, want_cls=True, want_attn=False)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"N={N} H={H} B={B} bias={'yes' if bias is not None else 'no'} kv={kv}: {us:.1f} us / call, "
          f"{4 * H * N * N * 64 * B / us / 1e6:.1f} TFLOP/s")

if not kv:
    # masked layers as the forward runs them: mask operands precomputed once per layer by cls_stat_mask, fetched by bulk copies
    from vision_transformer_cam_b200 import _lib
    aug = torch.zeros((B, int(_lib.load().vtc_attention_mask_operand_bytes(N))), dtype=torch.uint8, device=dev)
    cls_rows = torch.rand((B, H, N), device=dev).softmax(-1).contiguous()
    _, _, _, kb2, aug = ops.cls_stat_mask(cls_rows, 0.9, scale=0.125)
    for _ in range(3):
        ops.attention_masked(qkv, H, 0.125, kb2, aug)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention_masked(qkv, H, 0.125, kb2, aug)
    e1.record()
    torch.cuda.synchronize()
    print(f"N={N} H={H} B={B} bias=yes (precomputed mask operands, bg fraction {float((kb2 != 0).float().mean()):.2f}): {e0.elapsed_time(e1) * 100:.1f} us / call")
