"""profiles/<tag>_traffic.json from gpurun_out/traffic.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum over >= 2 bench steps): DRAM bytes per launch for every kernel family of ONE full step."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src_csv = sys.argv[2] if len(sys.argv) > 2 else "traffic.csv"
build = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", src_csv))))
hdr, data = None, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))


def tobytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


per = collections.OrderedDict()
for d in data:
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
    e = per.setdefault(d["ID"], {"name": name})
    m = d["Metric Name"]
    if m.startswith("dram__bytes"):
        e[m] = tobytes(d["Metric Value"], d["Metric Unit"])
    elif m.startswith("gpu__time"):
        v = float(d["Metric Value"].replace(",", ""))
        e["us"] = v / 1e3 if d["Metric Unit"] == "ns" else v
ids = list(per.keys())
st = [i for i, k in enumerate(ids) if per[k]["name"].startswith("stem_patchify")]
assert len(st) >= 1, "need at least one full step in the capture"
# two stems: the launches between them; one stem: the capture is exactly one forward (tools/r02b_ncu_step.sh)
step = [per[k] for k in (ids[st[0]:st[1]] if len(st) >= 2 else ids[st[0]:])]
agg = collections.defaultdict(lambda: {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "us": 0.0})
for k in step:
    fam = ("gemm (gemm_tc_kernel + gemm_pair_kernel + gemm_pair_x3_kernel + mlp_fused_kernel + mlp_fused_x3_kernel, all tcgen05 launches)"
           if k["name"].startswith(("gemm_tc_kernel", "gemm_pair_kernel", "gemm_pair_x3_kernel", "mlp_fused_kernel",
                                    "mlp_fused_x3_kernel")) else k["name"])
    a = agg[fam]
    a["launches"] += 1
    a["dram_read"] += k.get("dram__bytes_read.sum", 0)
    a["dram_write"] += k.get("dram__bytes_write.sum", 0)
    a["us"] += k.get("us", 0)
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
                 "one bench step (B=64), cold caches", "build": build, "step_kernels": len(step), "families": {}}
for f, a in agg.items():
    out["families"][f] = {"launches": a["launches"],
                          "dram_bytes_per_launch": (a["dram_read"] + a["dram_write"]) / a["launches"],
                          "dram_read_total": a["dram_read"], "dram_write_total": a["dram_write"], "us_total": a["us"]}
path = os.path.join(ROOT, "profiles", tag if tag.endswith(".json") else f"{tag}_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path, {f: round(v["dram_bytes_per_launch"] / 1e6, 1) for f, v in out["families"].items()})
