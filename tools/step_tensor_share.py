"""Time-weighted tensor-pipe activity of the forward from an ncu launch list (tools/r02b_ncu_step.sh):
sum_k(duration_k * tensor_pipe_active_k) / sum_k(duration_k) over every kernel of the captured window, plus the per-kernel table.
Usage: python tools/step_tensor_share.py gpurun_out/step_launches.csv [out.md]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, per = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            k = per.setdefault(d["ID"], {"name": d["Kernel Name"]})
            v = float(d["Metric Value"].replace(",", "") or 0)
            unit = d["Metric Unit"]
            if d["Metric Name"].startswith("gpu__time_duration"):
                v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)   # -> us
            if d["Metric Name"].startswith("dram__bytes"):
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            if d["Metric Name"].startswith("sm__cycles_elapsed"):
                v *= {"hz": 1, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9}.get(unit, 1) if unit.lower().endswith("hz") else 1
                v *= {"hz": 1, "khz": 1e3, "mhz": 1e6, "ghz": 1e9}.get(unit.lower(), 1) if False else 1
            k[d["Metric Name"]] = (v, unit)
    return list(per.values())


def short(name):
    return re.sub(r"\(.*", "", name).replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")


def main():
    launches = load(sys.argv[1])
    T = "gpu__time_duration.sum"
    E = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
    A = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    agg = collections.OrderedDict()
    tot_t = tot_w = tot_b = 0.0
    for L in launches:
        t = L[T][0]
        e = L.get(E, (0.0, ""))[0]
        a = L.get(A, (0.0, ""))[0]
        b = L.get("dram__bytes_read.sum", (0, ""))[0] + L.get("dram__bytes_write.sum", (0, ""))[0]
        g = agg.setdefault(short(L["name"]), [0, 0.0, 0.0, 0.0, 0.0])
        g[0] += 1; g[1] += t; g[2] += t * e; g[3] += t * a; g[4] += b
        tot_t += t; tot_w += t * e; tot_b += b
    lines = ["%d launches, %.1f us total (ncu per-launch times are cold-cache and serialised: compare SHARES)" % (len(launches), tot_t),
             "time-weighted tensor-pipe activity of the window (sm__pipe_tensor_cycles_active, %% of elapsed): %.1f %%" % (tot_w / tot_t),
             "tcgen05 kernels only: %.1f %%" % (sum(g[2] for k, g in agg.items() if g[2] > 0) / max(1e-9, sum(g[1] for k, g in agg.items() if g[2] > 0))),
             "DRAM bytes of the window: %.2f GB" % (tot_b / 1e9), "",
             "%10s %5s %6s %9s %8s %8s %9s  kernel" % ("us", "n", "share", "avg us", "tens%el", "tens%ac", "MB/launch")]
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%10.1f %5d %5.1f%% %9.1f %8.1f %8.1f %9.1f  %s" % (g[1], g[0], 100 * g[1] / tot_t, g[1] / g[0], g[2] / g[1], g[3] / g[1], g[4] / g[0] / 1e6, k))
    txt = "\n".join(lines)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write("```\n" + txt + "\n```\n")
    if len(sys.argv) > 3:   # machine-readable digest for bench.py's roofline record
        import json
        tc_t = sum(g[1] for g in agg.values() if g[2] > 0)
        json.dump({"launches": len(launches), "window_us": tot_t,
                   "tensor_pipe_pct_of_elapsed_forward": tot_w / tot_t,
                   "tensor_pipe_pct_of_elapsed_tcgen05_kernels": sum(g[2] for g in agg.values()) / tc_t,
                   "tensor_pipe_pct_of_active_tcgen05_kernels": sum(g[3] for g in agg.values()) / tc_t,
                   "tcgen05_time_share": tc_t / tot_t, "dram_gb": tot_b / 1e9,
                   "build": sys.argv[4] if len(sys.argv) > 4 else None,
                   "how": "ncu --clock-control none, one parity-mode forward at B = 64 (tools/r02b_ncu_step.sh); per-launch "
                          "times are cold-cache and serialised"}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
