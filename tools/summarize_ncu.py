"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
  python tools/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.txt
  python tools/summarize_ncu.py zoo      gpurun_out/zoo_raw.csv [traffic.json]  > profiles/rNN_kernels_ncu.txt
       (zoo_raw.csv = `ncu -i zoo.ncu-rep --page raw --csv` of a tools/kernel_zoo.py capture)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_sample_count"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        k = row["Kernel Name"][:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total")
    print(f"# (cold-cache, serialised launches: compare SHARES with bench.py's CUDA-event numbers, not absolutes)")
    print(f"{'total ms':>10} {'share':>6} {'count':>6} {'avg us':>9}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t / 1e3:10.2f} {100 * t / tot:5.1f}% {c:6d} {t / c:9.1f}  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>16s} {units[i]}")


HBM_PEAK_GBS = 6545.9       # MEASURED_PEAKS.json hbm_gbs


def _num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def zoo(path, traffic_out=None):
    """One line per distinct kernel (template instantiation) of a `--set full` capture: launches, mean duration, DRAM bytes per
    launch, achieved DRAM GB/s and its fraction of the measured HBM peak, tensor-pipe activity, registers, and the top stalls."""
    import json
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in hdr}
    unit = {k: units[hdr.index(k)] for k in hdr}

    def get(r, k):
        if k not in col:
            return float("nan")
        v, u = _num(r[col[k]]), unit[k]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        return v * scale

    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        key = name.split("(")[0][:90]
        a = agg.setdefault(key, {"n": 0, "us": [], "rd": [], "wr": [], "tensor": [], "xu": [], "regs": None, "grid": grid, "issue": [], "l2": [],
                                 "stalls": collections.Counter(), "samples": 0})
        a["n"] += 1
        a["us"].append(get(r, "gpu__time_duration.sum"))
        a["rd"].append(get(r, "dram__bytes_read.sum"))
        a["wr"].append(get(r, "dram__bytes_write.sum"))
        a["tensor"].append(get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
        a["xu"].append(get(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"))
        a["issue"].append(get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"))
        a["l2"].append(get(r, "lts__t_bytes.sum"))
        a["regs"] = r[col["launch__registers_per_thread"]] if "launch__registers_per_thread" in col else "?"
        for k in col:
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                a["stalls"][k.replace("smsp__pcsamp_warps_issue_stalled_", "")] += _num(r[col[k]]) if r[col[k]] else 0
        a["samples"] += get(r, "smsp__pcsamp_sample_count") if "smsp__pcsamp_sample_count" in col else 0
    mean = lambda v: sum(v) / len(v) if v else float("nan")
    print(f"# ncu --set full --clock-control none, tools/kernel_zoo.py (ViT-B/16-224, B = 256; cold-cache single launches).  DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration;")
    print(f"# fraction of the measured HBM peak ({HBM_PEAK_GBS} GB/s, MEASURED_PEAKS.json).  tensor% = sm__pipe_tensor_cycles_active, xu% = MUFU pipe, issue% = smsp__issue_active.")
    print(f"{'kernel':72s} {'n':>3s} {'avg us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'of HBM':>6s} {'tensor%':>7s} {'xu%':>5s} {'issue%':>6s} {'regs':>4s}  top stalls")
    traffic = {}
    for k, a in agg.items():
        us, rd, wr = mean(a["us"]), mean(a["rd"]), mean(a["wr"])
        gbs = (rd + wr) / us / 1e3 if us > 0 else float("nan")
        top = ", ".join(f"{n} {100 * c / max(sum(a['stalls'].values()), 1):.0f}%" for n, c in a["stalls"].most_common(3))
        print(f"{k[:72]:72s} {a['n']:3d} {us:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {gbs:7.0f} {gbs / HBM_PEAK_GBS:6.2f} {mean(a['tensor']):7.1f} {mean(a['xu']):5.1f} {mean(a['issue']):6.1f} {a['regs']:>4s}  {top}")
        traffic[k] = {"launches": a["n"], "us": us, "dram_read": rd, "dram_write": wr, "l2_bytes": mean(a["l2"])}
    # the four GEMM kinds of a layer, by template epilogue and launch order (proj and fc2 share <2>: proj is the shorter one)
    if traffic_out:
        out = {}
        seq = [(r[col["Kernel Name"]], get(r, "gpu__time_duration.sum"), get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")) for r in rows[2:]]
        kinds = {"gemm_qkv": [], "gemm_fc1": [], "gemm_proj": [], "gemm_fc2": [], "gemm_patch": []}
        resid = [s for s in seq if "gemm2_bf16_kernel<2" in s[0] or "gemm2_bf16_kernel<(int)2" in s[0]]
        med = sorted(s[1] for s in resid)[len(resid) // 2] if resid else 0
        for s in seq:
            n = s[0]
            if "gemm2_bf16_kernel<0" in n or "gemm2_bf16_kernel<(int)0" in n: kinds["gemm_qkv"].append(s)
            elif "gemm2_bf16_kernel<1" in n or "gemm2_bf16_kernel<(int)1" in n: kinds["gemm_fc1"].append(s)
            elif "gemm2_bf16_kernel<5" in n or "gemm2_bf16_kernel<(int)5" in n: kinds["gemm_patch"].append(s)
            elif s in resid: kinds["gemm_proj" if s[1] < med else "gemm_fc2"].append(s)
        for k, v in kinds.items():
            if v:
                out[k] = {"launches": len(v), "us": mean([s[1] for s in v]), "dram_read": mean([s[2] for s in v]), "dram_write": mean([s[3] for s in v])}
        out["source"] = "ncu --set full --clock-control none on tools/kernel_zoo.py (B = 256, one fused forward of 6 layers), per-launch means"
        out["kernels"] = traffic
        json.dump(out, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    fn = {"launches": launches, "kernel": kernel, "zoo": zoo}[sys.argv[1]]
    fn(*sys.argv[2:])
