#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "gemm_x3 or yolox or north_star or b64" 2>&1 | tail -5
for m in 3 5; do echo "== GELU mode $m"; GDRN_X3_GELU_MODE=$m timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "'shape'" | grep fc1 | cut -c1-120; done
echo "== trace"; GDRN_GEMM_TRACE=1 timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "trace" | grep "epi=1" | sort | uniq -c | sort -rn | awk '{ if (!seen[$6$7$8]++) print }' | head -6 | cut -c1-420
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/bench_step14.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_step14.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline'].get('frac'), d.get('breakdown_ms'))
PY
