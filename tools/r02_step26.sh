#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "forward_vs_oracle or north_star or predictor or sharded" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/bench_step26.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_step26.json'))
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('gpu_launches'))
PY
