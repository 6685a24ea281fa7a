"""Summarise ncu reports (gpurun_out/*.ncu-rep) and launch lists into profiles/ (tracked)."""
import csv
import collections
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def rep_summary(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = ["kernel: " + d.get("Kernel Name", ("?", ""))[0]]
    for k in KEYS:
        if k in d:
            lines.append(f"  {k} = {d[k][0]} {d[k][1]}")
    return "\n".join(lines)


def launch_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        if not d["Metric Name"].startswith("gpu__time_duration"):
            continue
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Unit"] == "ns":
            v /= 1e3
        k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"{sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{t:10.1f} us {n:5d}x {100 * t / tot:5.1f}%  avg {t / n:8.1f} us  {k}")
    return "\n".join(lines)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    go = os.path.join(ROOT, "gpurun_out")
    out = []
    lp = os.path.join(go, "launches.csv")
    if os.path.exists(lp):
        out.append("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3)\n```\n"
                   + launch_summary(lp) + "\n```\n")
    for f in sorted(os.listdir(go)):
        if f.endswith(".ncu-rep"):
            out.append(f"## {f} (ncu --set full --clock-control none)\n```\n" + rep_summary(os.path.join(go, f)) + "\n```\n")
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md")
    open(path, "w").write("\n".join(out))
    print("wrote", path)
