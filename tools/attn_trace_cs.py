"""Timeline of the column-split attention kernel (debug): per (CTA, step, warp) %clock64 stamps (SM clocks).
softmax warps 0-15, slots: 0 step top, 1 S ready, 2 chunks done, 3 row barrier passed, 4 p_full arrived, 5 O ready, 6 item done
MMA threads (rows 16, 17), slots: 0 step top, 1 O region free, 2 K landed (QK issued), 3 QK committed, 4 P ready, 5 V landed (PV issued)

The stamps are compiled in only with -DVTC_ACS_TRACE=1:
    tools/build_ablate.sh 10000000 && VTC_LIB_PATH=$PWD/tools/ab/libvtc_ablate10000000.so python tools/attn_trace_cs.py"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_transformer_cam_b200 import _lib, ops
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 197
H = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
WANT_CLS = (sys.argv[4] != "nocls") if len(sys.argv) > 4 else True
dev = torch.device("cuda:0")
qkv = torch.randn((B, N, 3 * H * 64), device=dev).bfloat16()
for _ in range(3):
    ops.attention(qkv, H, 0.125, want_cls=WANT_CLS)
trace = torch.zeros((148, 64, 18, 8), dtype=torch.int64, device=dev)
lib.vtc_debug_set_attention_trace(ctypes.c_void_p(trace.data_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(qkv, H, 0.125, want_cls=WANT_CLS); e1.record(); torch.cuda.synchronize()
lib.vtc_debug_set_attention_trace(ctypes.c_void_p(0))
print("kernel ms", e0.elapsed_time(e1))
t = trace.cpu().double()
cta = 5
t0 = t[cta, 0, 0, 0]
sm_names = ["waitS", "chunks", "rowbar", "arrive", "waitO", "epi"]
mma_names = ["waitOfree", "waitK", "issueQK", "waitP", "waitV"]
for s in range(3, 8):
    for w in (0, 4, 8, 12):
        r = t[cta, s, w]
        print(f"  step {s} warp {w:2d}: start {float(r[0]-t0):9.0f} clk | " + " ".join(f"{n} {float(r[k+1]-r[k]):6.0f}" for k, n in enumerate(sm_names)))
    for w in (16, 17):
        r = t[cta, s, w]
        print(f"  step {s} MMA  {w-16}: start {float(r[0]-t0):9.0f} clk | " + " ".join(f"{n} {float(r[k+1]-r[k]):6.0f}" for k, n in enumerate(mma_names)))
d = t[:, 4:18]
for w in (0, 1, 2, 4, 5, 6, 8, 12):
    print(f"warp {w:2d} mean (clk): " + "  ".join(f"{n} {float((d[:, :, w, k+1]-d[:, :, w, k]).mean()):.0f}" for k, n in enumerate(sm_names)),
          " period", float((d[:, 1:, w, 0] - d[:, :-1, w, 0]).mean()))
for w in (16, 17):
    print(f"MMA {w-16} mean (clk): " + "  ".join(f"{n} {float((d[:, :, w, k+1]-d[:, :, w, k]).mean()):.0f}" for k, n in enumerate(mma_names)),
          " period", float((d[:, 1:, w, 0] - d[:, :-1, w, 0]).mean()))
print("---- absolute timeline of CTA %d (clk since step 0): softmax slots S_ready / chunks_done / p_arrive / O_ready / item_done; MMA slots QK_issued / P_ready / PV_issued" % cta)
for s in range(4, 10):
    for w, name in ((0, "g0 part0"), (4, "g0 part1"), (8, "g1 part0"), (12, "g1 part1")):
        r = t[cta, s, w] - t0
        print(f"  step {s} {name}: top {float(r[0]):8.0f}  S_ready {float(r[1]):8.0f}  chunks_done {float(r[2]):8.0f}  p_arrive {float(r[4]):8.0f}  O_ready {float(r[5]):8.0f}  done {float(r[6]):8.0f}")
    for w in (16, 17):
        r = t[cta, s, w] - t0
        print(f"  step {s} MMA g{w-16}  : top {float(r[0]):8.0f}  O_free {float(r[1]):8.0f}  QK_issue {float(r[2]):8.0f}  QK_done_issue {float(r[3]):8.0f}  P_ready {float(r[4]):8.0f}  PV_issue {float(r[5]):8.0f}  | producer: K_load_issued {float(r[6]):8.0f}  V_load_issued {float(r[7]):8.0f}")
