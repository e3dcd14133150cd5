"""Timeline of the persistent attention kernel (debug): per (CTA, item, softmax group) %globaltimer stamps.
slots: 0 loop top, 1 S ready, 2 P written, 3 CLS row written, 4 O ready, 5 O stored"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_transformer_cam_b200 import _lib, ops
lib = _lib.load()
B, N, H = 256, 197, 12
dev = torch.device("cuda:0")
qkv = torch.randn((B, N, 3 * H * 64), device=dev).bfloat16()
for _ in range(3):
    ops.attention(qkv, H, 0.125)
trace = torch.zeros((148, 32, 2, 8), dtype=torch.int64, device=dev)
lib.vtc_debug_set_attention_trace(ctypes.c_void_p(trace.data_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(qkv, H, 0.125); e1.record(); torch.cuda.synchronize()
lib.vtc_debug_set_attention_trace(ctypes.c_void_p(0))
print("kernel ms", e0.elapsed_time(e1))
t = trace.cpu().double()
names = ["wait S", "softmax", "cls", "wait O", "store"]
for cta in (0, 77):
    t0 = t[cta, 0, 0, 0]
    print(f"CTA {cta}: per item, group: start(us) | " + " ".join(names) + " (ns)")
    for i in range(2, 10):
        for g in range(2):
            r = t[cta, i, g]
            print(f"  item {i} g{g}: {float(r[0]-t0)/1e3:8.2f} | " + " ".join(f"{float(r[k+1]-r[k]):7.0f}" for k in range(5)))
d = t[:, 4:20]
for g in range(2):
    print(f"group {g} mean (ns): " + "  ".join(f"{n} {float((d[:, :, g, k+1]-d[:, :, g, k]).mean()):.0f}" for k, n in enumerate(names)),
          " period", float((d[:, 1:, g, 0] - d[:, :-1, g, 0]).mean()))
