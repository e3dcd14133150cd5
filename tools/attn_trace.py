"""Phase timeline of the attention kernel (debug): %globaltimer stamps per CTA.
slots: 0 start, 1 after prologue sync, 6 Q/K landed (control thread), 2 S ready (softmax warp 0), 3 P written, 4 O ready, 5 epilogue done, 7 exit"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_transformer_cam_b200 import _lib, ops
lib = _lib.load()
B, N, H = 256, 197, 12
dev = torch.device("cuda:0")
qkv = (torch.randn((B, N, 3 * H * 64), device=dev) * 1.0).bfloat16()
for _ in range(3):
    ops.attention(qkv, H, 0.125)
nt = (N + 127) // 128
trace = torch.zeros((B * H * nt, 8), dtype=torch.int64, device=dev)
lib.vtc_debug_set_attention_trace(ctypes.c_void_p(trace.data_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(qkv, H, 0.125); e1.record(); torch.cuda.synchronize()
lib.vtc_debug_set_attention_trace(ctypes.c_void_p(0))
t = trace.cpu().double()
print("kernel ms", e0.elapsed_time(e1))
t0 = t[:, 0].min()
names = [("prologue", 0, 1), ("load wait (ctl)", 1, 6), ("QK mma->S ready", 6, 2), ("softmax", 2, 3), ("PV + O wait", 3, 4), ("epilogue", 4, 5), ("teardown", 5, 7), ("total", 0, 7)]
for mt in range(nt):
    sel = t[mt::nt]
    print(f"m-tile {mt}: " + "  ".join(f"{n} {float((sel[:, b] - sel[:, a]).mean()):.0f}ns" for n, a, b in names))
print("span of all CTAs (us):", float(t[:, 7].max() - t0) / 1e3, " sum CTA lifetimes / (296 slots) us:", float((t[:, 7] - t[:, 0]).sum()) / 296 / 1e3)
