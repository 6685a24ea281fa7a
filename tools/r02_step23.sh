#!/bin/bash
cd "$(dirname "$0")/.."
for t in 1 0 1 0; do echo "TOKEN=$t"; GDRN_DW_TOKEN=$t timeout 200 python tools/dw_bench.py 2>&1 | grep "variant=1" | grep "split=1"; done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/bench_step23.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_step23.json'))
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('gpu_launches'))
PY
