#!/bin/bash
# final evidence of round 2 (second session), under gpurun: full GPU suite, smoke, the driver's bench command + reference arm,
# ncu --set full of the new FC kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$(date +%s)
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench_final.json
t1=$(date +%s); echo "bench wall $((t1-t0)) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json')); r=d['roofline']
print('bench', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), d['clocks'], 'frac', r['frac'], 'exec', r.get('tensor_executed_frac'), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
print('gemm', r['gemm_ms_per_step'], 'dw', r['dwconv_ms_per_step'], 'other', r['other_ms_per_step'], 'launches', d.get('gpu_launches'), 'tensor_pipe', r.get('tensor_pipe'))
print('alt', d.get('bf16_mode') and d['bf16_mode']['value']); print('backbones', [(b['arch'], round(b['value'])) for b in d.get('other_backbones') or []])
print('eager', json.dumps(d.get('gpu_eager_baseline'))[:600])
for n in d.get('native_ops') or []:
    print('  ', n.get('op'), n.get('ms'), n.get('frac'), n.get('rois_per_s'), n.get('error'))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_arm.json; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
export GDRN_BENCH_GRAPHS=0
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fc_f32_partial_kernel' -s 3 -c 1 \
  -o gpurun_out/prof_fc_partial -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_fc.log 2>&1
tail -1 gpurun_out/ncu_fc.log | cut -c1-120
