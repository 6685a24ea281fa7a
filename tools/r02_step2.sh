#!/bin/bash
# round-2 step 2: full GPU suite + smoke + the rewritten bench (parity mode of record, secondary records)
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -25 > gpurun_out/step2_tests.log
cat gpurun_out/step2_tests.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/step2_smoke.log 2>&1; tail -2 gpurun_out/step2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_step2.json 2> gpurun_out/bench_step2.err; tail -3 gpurun_out/bench_step2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_step2.json')); r=d['roofline']
print('value %.1f e2e %.1f ms %.3f gemm %.3f dw %.3f other %.3f frac %.3f exec_frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['gemm_ms_per_step'], r['dwconv_ms_per_step'], r['other_ms_per_step'], r['frac'], r.get('tensor_executed_frac',0)))
print('bf16_mode', d.get('bf16_mode'))
print('eager', d.get('gpu_eager_baseline'))
print('cpu', d.get('cpu_baseline'))
for n in d.get('native_ops') or []:
    print({k: (round(v,4) if isinstance(v,float) else v) for k,v in n.items() if k in ('op','ms','achieved','unit','frac','reference_cuda_ms','error','rois_per_s','images_per_s','cpu_reference_ms_per_cloud','cpu_port_ms_per_problem')})
PY
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
