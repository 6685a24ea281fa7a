#!/bin/bash
cd "$(dirname "$0")/.."
echo "== dw no token"; timeout 200 python tools/dw_bench.py 2>&1 | grep "variant=1"
echo "== dw token"; GDRN_DW_TOKEN=1 timeout 200 python tools/dw_bench.py 2>&1 | grep "variant=1"
echo "== x3 shapes (A&S gelu, swizzle64 staging)"; timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "fc1"
timeout 900 python -m pytest tests/ -x -q -m gpu -k "not sharded" 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_step8.json 2> gpurun_out/bench_step8.err; tail -2 gpurun_out/bench_step8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_step8.json')); r=d['roofline']
print('value %.1f e2e %.1f ms %.3f gemm %.3f dw %.3f other %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['gemm_ms_per_step'], r['dwconv_ms_per_step'], r['other_ms_per_step']))
PY
timeout 300 python bench.py --workload ycbv5 2>&1 | tail -2 | cut -c1-400
