#!/bin/bash
# same-box A/B of two library builds (under gpurun): tools/ab/libgdrn_b200_prev.so vs the in-tree build, interleaved
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp gdrnpp_bop2022_b200/libgdrn_b200.so /tmp/cur.so
run() {
  cp "$1" gdrnpp_bop2022_b200/libgdrn_b200.so
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/ab.json
  python - "$2" <<'PY'
import json, sys
d = json.load(open('gpurun_out/ab.json'))
print(sys.argv[1], round(d['value'], 1), 'ROIs/s', round(d['ms_per_step'], 3), 'ms', d['clocks']['sm_mhz'], 'MHz', d.get('gpu_launches'))
PY
}
for i in 1 2 3; do run tools/ab/libgdrn_b200_prev.so prev; run /tmp/cur.so cur; done
cp /tmp/cur.so gdrnpp_bop2022_b200/libgdrn_b200.so
