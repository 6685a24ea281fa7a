#!/bin/bash
# session-2 validation of HEAD (under gpurun): full GPU suite, smoke, the driver's bench line (wall-clock timed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests -q -x -m gpu --durations=8 2>&1 | tail -16
t1=$(date +%s); echo "pytest wall $((t1-t0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
t2=$(date +%s); echo "smoke wall $((t2-t1)) s"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_call1.err | tail -1 > gpurun_out/bench_call1.json
t3=$(date +%s); echo "bench wall $((t3-t2)) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_call1.json'))
r=d['roofline']
print('bench', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), d['clocks'], 'frac', r['frac'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
print('gemm', r['gemm_ms_per_step'], 'dw', r['dwconv_ms_per_step'], 'other', r['other_ms_per_step'], 'launches', d.get('gpu_launches'))
print('alt', d.get('bf16_mode')); print('backbones', d.get('other_backbones'))
PY
tail -5 gpurun_out/bench_call1.err
