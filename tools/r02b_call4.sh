#!/bin/bash
# split-K FC stack: parity tests that run the FC stack (all batch sizes), then the one-step ncu launch list and an A/B-free bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "forward or north_star or b64 or sharded or precisions or predictor or deterministic" 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > gpurun_out/bench_call4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_call4.json')); r=d['roofline']
print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], 'gemm', round(r['gemm_ms_per_step'],3), 'dw', round(r['dwconv_ms_per_step'],3), 'other', round(r['other_ms_per_step'],3), 'launches', d.get('gpu_launches'))
PY
tools/r02b_ncu_step.sh
