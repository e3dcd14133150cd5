"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): tcgen05.mma = UTC*MMA,
tcgen05.ld / st = LDTM / STTM, TMA = UTMALDG / UTMASTG / UTMAREDG / UBLKCP (cp.async.bulk), legacy warp-level tensor path =
HMMA, plus registers / spills from `cuobjdump -res-usage`.  Runs on the CPU box:

    python tools/sass_summary.py [path/to/libvtc.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vision_transformer_cam_b200", "libvtc.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2", "LDGSTS", "RED", "ATOM"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in MNEMONICS:
            if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                counts[cur][k] += 1
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
fn = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and fn:
        usage[fn] = tuple(int(v) for v in m.groups())
arch = re.findall(r"\.(sm_\w+)\.cubin", subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout)
print(f"# {os.path.basename(LIB)}: SASS mnemonic counts per kernel (cuobjdump -sass), registers / stack / static smem (cuobjdump -res-usage)")
print(f"# embedded cubins: {sorted(set(arch))}")
hdr = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "MUFU.EX2", "FFMA2"]
print(f"{'kernel':100s} {'instr':>6s} {'REG':>4s} {'STACK':>5s} " + " ".join(f"{h:>8s}" for h in hdr))
for name, c in sorted(counts.items(), key=lambda kv: demangle(kv[0])):
    d = re.sub(r"\(.*", "", demangle(name)).replace("vtc::", "").replace("(anonymous namespace)::", "")
    r = usage.get(name, (0, 0, 0))
    print(f"{d[:100]:100s} {c['_total']:6d} {r[0]:4d} {r[1]:5d} " + " ".join(f"{c[h]:8d}" for h in hdr))
tc = [demangle(n) for n, c in counts.items() if c["UTCHMMA"]]
hm = [demangle(n) for n, c in counts.items() if c["HMMA"]]
print(f"# kernels with tcgen05.mma (UTCHMMA): {len(tc)}; kernels with legacy HMMA: {len(hm)}")
for n in hm:
    print("#   HMMA:", re.sub(r"\(.*", "", n))
