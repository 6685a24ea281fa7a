#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -5
for pdl in 0 1 0 1; do
  GDRN_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/bench_step15_pdl$pdl.json
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_step15_pdl$pdl.json'))
print('PDL=$pdl', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
