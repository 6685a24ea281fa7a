"""Summarise an ncu source page (ncu -i X.ncu-rep --page source --csv): top SASS instructions by stall samples."""
import csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in data if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, " instructions", len(data))
for s in stalls:
    v = sum(int(r[ix[s]]) for r in data if r[ix[s]].isdigit())
    if v * 50 > tot:
        print("  %-28s %6d  %.1f%%" % (s, v, 100.0 * v / tot))
for n, r in enumerate(sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]):
    st = {s: int(r[ix[s]]) for s in stalls if r[ix[s]].isdigit() and int(r[ix[s]]) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(8), r[ix["Source"]][:64].ljust(64),
          " ".join("%s=%d" % (k[6:], v) for k, v in st))
