#!/bin/bash
# round-2 final evidence run (under gpurun): full GPU suite, smoke, the driver's bench line + reference arm, launch list + ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python bench.py 2>&1 | tail -1 > gpurun_out/r02_bench_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('bench', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), d['clocks'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
print('alt', d.get('bf16_mode')); print('backbones', d.get('other_backbones')); print('eager', json.dumps(d.get('gpu_eager_baseline'))[:400])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r02_bench_reference_arm.json; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
