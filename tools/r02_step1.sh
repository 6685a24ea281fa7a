#!/bin/bash
# round-2 step 1: correctness of the CTA-pair split-bf16 kernel + parity-mode model tests + parity-mode bench
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm_x3 or north_star or b64 or sharded or precisions_agree" 2>&1 | tail -15 > gpurun_out/step1_tests.log
cat gpurun_out/step1_tests.log
timeout 300 python bench.py --precision bf16x3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_x3_step1.json 2> gpurun_out/bench_x3_step1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_x3_step1.json')); r=d['roofline']
print('x3 value %.1f ms %.3f gemm %.3f dw %.3f other %.3f' % (d['value'], d['ms_per_step'], r['gemm_ms_per_step'], r['dwconv_ms_per_step'], r['other_ms_per_step']))
PY
