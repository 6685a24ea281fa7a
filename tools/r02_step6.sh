#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dwconv_ln_pingpong" 2>&1 | tail -12
echo "== traces"; GDRN_GEMM_TRACE=1 timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "trace" | sort | uniq -c | sort -rn | awk '{ if (!seen[$6$7$8$9]++) print }' | head -12
for pp in 0 1; do
GDRN_DW_PP=$pp timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_step6_pp$pp.json 2> gpurun_out/bench_step6_pp$pp.err; tail -2 gpurun_out/bench_step6_pp$pp.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_step6_pp$pp.json')); r=d['roofline']
print('pp=$pp value %.1f e2e %.1f ms %.3f gemm %.3f dw %.3f other %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['gemm_ms_per_step'], r['dwconv_ms_per_step'], r['other_ms_per_step']))
PY
done
