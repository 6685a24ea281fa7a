#!/bin/bash
# last evidence refresh (under gpurun): full GPU suite, smoke, the driver's bench command, one-step ncu launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json')); r=d['roofline']
print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'frac', round(r['frac'],4), 'exec', round(r.get('tensor_executed_frac'),4))
print('gemm', round(r['gemm_ms_per_step'],3), 'dw', round(r['dwconv_ms_per_step'],3), 'other', round(r['other_ms_per_step'],3), 'launches', d.get('gpu_launches'))
print('alt', d.get('bf16_mode') and round(d['bf16_mode']['value'])); print('backbones', [(b['arch'], round(b['value'])) for b in d.get('other_backbones') or []])
for n in (d.get('native_ops') or [])[-2:]:
    print('  ', n.get('op')[:40], n.get('ms'), n.get('rois_per_s'), n.get('error'))
PY
tools/r02b_ncu_step.sh | tail -36
