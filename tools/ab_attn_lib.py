"""Time vtc_attention of two builds of libvtc on the same box (debug): python tools/ab_attn_lib.py libA.so libB.so"""
import ctypes, os, sys, torch
N, H, B = (int(v) for v in os.environ.get("VTC_AB_SHAPE", "197,12,256").split(","))
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn((B, N, 3 * H * 64), generator=g, device=dev).bfloat16()
kb = torch.where(torch.rand((B, N), generator=g, device=dev) < 0.3, -100.0, 0.0)
kb[:, 0] = 0
out = torch.empty((B, N, H * 64), dtype=torch.bfloat16, device=dev)
cls = torch.empty((B, H, N), device=dev)
libs = [(p, ctypes.CDLL(p)) for p in sys.argv[1:]]
P, I, F = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float
for _, lib in libs:
    lib.vtc_attention.argtypes = [P, P, P, P, P, I, I, I, F, P]
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    for path, lib in libs:
        for bias in (None, kb):
            call = lambda: lib.vtc_attention(qkv.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(), cls.data_ptr(), None, B, N, H, 0.125, st)
            for _ in range(3):
                assert call() == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                call()
            e1.record()
            torch.cuda.synchronize()
            print(f"{path.split('/')[-1]:24s} bias={'yes' if bias is not None else 'no '}: {e0.elapsed_time(e1) * 100:.1f} us")
outs = []
for path, lib in libs:          # how far apart the builds' results are (max |difference| of O and of the CLS rows, against the first library)
    assert lib.vtc_attention(qkv.data_ptr(), kb.data_ptr(), out.data_ptr(), cls.data_ptr(), None, B, N, H, 0.125, st) == 0
    torch.cuda.synchronize()
    outs.append((out.float().clone(), cls.clone()))
    print(f"{path.split('/')[-1]:24s} vs first: O {float((outs[-1][0] - outs[0][0]).abs().max()):.2e}  cls rows {float((outs[-1][1] - outs[0][1]).abs().max()):.2e}")
if os.environ.get("VTC_AB_PLAIN_ONLY") == "1":
    sys.exit(0)
print("---- vtc_attention_mean (attention + packed P + head mean)")
Z = ctypes.c_size_t
mean = torch.empty((B, N, N), device=dev)
for path, lib in libs:
    lib.vtc_attention_mean_scratch_bytes.restype = Z
    lib.vtc_attention_mean_scratch_bytes.argtypes = [I, I, I]
    lib.vtc_attention_mean.argtypes = [P, P, P, P, P, P, Z, I, I, I, F, P]
scr = {path: torch.empty((lib.vtc_attention_mean_scratch_bytes(B, N, H),), dtype=torch.uint8, device=dev) for path, lib in libs}
for rep in range(3):
    for path, lib in libs:
        call = lambda: lib.vtc_attention_mean(qkv.data_ptr(), None, out.data_ptr(), cls.data_ptr(), mean.data_ptr(), scr[path].data_ptr(), scr[path].numel(),
                                              B, N, H, 0.125, st)
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        print(f"{path.split('/')[-1]:24s} attention_mean: {e0.elapsed_time(e1) * 100:.1f} us   rowsum err {float((mean.sum(-1) - 1).abs().max()):.1e}")
