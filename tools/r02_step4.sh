#!/bin/bash
# round-2 step 4: branch-free erf epilogue, cluster FPS, mesh registry -- tests + parity-mode bench
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/ -x -q -m gpu -k "not b64 and not sharded" 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_step4.json 2> gpurun_out/bench_step4.err; tail -2 gpurun_out/bench_step4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_step4.json')); r=d['roofline']
print('value %.1f e2e %.1f ms %.3f gemm %.3f dw %.3f other %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['gemm_ms_per_step'], r['dwconv_ms_per_step'], r['other_ms_per_step']))
PY
