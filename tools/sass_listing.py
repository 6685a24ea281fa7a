"""profiles/<tag>_sass_tcgen05_kernels.txt: SASS mnemonic counts (cuobjdump -sass) of every tcgen05 / TMA / depthwise kernel of the
built library, as evidence that the hot path is hand-written Blackwell code (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,
UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, UTCBAR = tcgen05.commit, FFMA2 / FMUL2 / FADD2 = packed fp32).
Usage: python tools/sass_listing.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "gdrnpp_bop2022_b200", "libgdrn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = re.compile(r"^(UTC|LDTM|STTM|UTMA|UBLKCP|FFMA2|FMUL2|FADD2|MUFU|UTCBAR|SYNCS\.ARRIVE\.TRANS|ACQBULK|UCGABAR)")
KERNELS = re.compile(r"gemm_|mlp_fused|dwconv|ln_patchify|gn_gelu|fc_f32")
out = ["# SASS evidence for the tcgen05 / TMA / packed-fp32 kernels of libgdrn_b200.so (cuobjdump -sass, sm_100a; %s build)" % tag,
       "# per kernel: instruction count and Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG /",
       "# UTMAREDG = TMA load / store / reduce, UTCBAR = tcgen05.commit, FFMA2 / FMUL2 / FADD2 = packed fp32).", ""]
cur, counts, n = None, None, 0


def flush():
    if cur and KERNELS.search(cur):
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip() or cur
        items = "  ".join("%s x%d" % kv for kv in sorted(counts.items()))
        out.append("## " + name)
        out.append("instructions: %d   %s" % (n, items))
        out.append("")


for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        n += 1
        op = m.group(1)
        if WANT.match(op):
            key = op
            if op.startswith("MUFU"):
                key = ".".join(op.split(".")[:2])
            elif op.startswith(("FFMA2", "FMUL2", "FADD2")):
                key = op.split(".")[0]
            counts[key] += 1
flush()
path = os.path.join(ROOT, "profiles", "%s_sass_tcgen05_kernels.txt" % tag)
open(path, "w").write("\n".join(out))
print("wrote", path, len(out) // 3, "kernels")
