#!/bin/bash
# batch-aware tile widths: GEMM + forward parity tests, small-batch forward times, per-image workload, B = 64 headline check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm_x3_vs or forward or north_star or b64 or tiny_small or predictor or deterministic" 2>&1 | tail -4
for b in 1 5 16 32; do timeout 300 python tools/fwd_small.py $b 3 2>&1 | tail -1; done
timeout 300 python bench.py --workload ycbv5 2>/dev/null | tail -1 | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > gpurun_out/bench_call6.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_call6.json')); r=d['roofline']
print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], 'gemm', round(r['gemm_ms_per_step'],3), 'dw', round(r['dwconv_ms_per_step'],3), 'other', round(r['other_ms_per_step'],3), 'launches', d.get('gpu_launches'))
PY
