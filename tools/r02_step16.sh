#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "mlp_fused_x3" 2>&1 | tail -8
echo "== trace"; GDRN_MLP_TRACE=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "mlp_fused_x3" 2>&1 | grep "trace" | head -3 | cut -c1-500
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "north_star or b64 or sharded" 2>&1 | tail -4
for f in 0 1 0 1; do
  GDRN_MLP_FUSED_X3=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/bench_step16_f$f.json
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_step16_f$f.json'))
print('FUSED_X3=$f', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('gpu_launches'))
PY
done
