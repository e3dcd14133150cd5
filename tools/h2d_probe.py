"""Host-link probe (debug): the process's cpuset, the GPU's NVML-ideal CPUs, the NUMA nodes, and the pinned-memory H2D / D2H bandwidth
with the allocating thread bound to each NUMA node in turn.    python tools/h2d_probe.py"""
import glob, os, time
import torch
import pynvml

dev = torch.device("cuda:0")
torch.cuda.init()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
cpus = sorted(os.sched_getaffinity(0))
print("cpuset of the process:", cpus)
try:
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    ideal = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    print("NVML ideal CPUs of GPU 0:", ideal[:8], "...", len(ideal), "cpus; intersection with the cpuset:", sorted(set(ideal) & set(cpus)))
except Exception as e:
    ideal = []
    print("nvmlDeviceGetCpuAffinity failed:", e)
nodes = {}
for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    txt = open(p).read().strip()
    ids = []
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-"); ids += list(range(int(a), int(b) + 1))
        elif part:
            ids.append(int(part))
    nodes[p.split("/")[-2]] = ids
    print(p.split("/")[-2], txt, "-> in cpuset:", len(set(ids) & set(cpus)))


def bw(tag):
    x = torch.empty((256, 3, 224, 224)).pin_memory()
    x.normal_()
    d = torch.empty_like(x, device=dev)
    for _ in range(2):
        d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(10):
        x.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    gb = x.numel() * 4 * 10 / 1e9
    print(f"{tag:40s} H2D {gb / (t1 - t0):6.1f} GB/s   D2H {gb / (t2 - t1):6.1f} GB/s", flush=True)


bw("default affinity")
for name, ids in nodes.items():
    use = sorted(set(ids) & set(cpus))
    if not use:
        continue
    os.sched_setaffinity(0, use)
    bw(f"allocating thread on {name} ({len(use)} cpus)")
os.sched_setaffinity(0, cpus)
if ideal and set(ideal) & set(cpus):
    os.sched_setaffinity(0, sorted(set(ideal) & set(cpus)))
    bw("NVML ideal CPUs")
    os.sched_setaffinity(0, cpus)
