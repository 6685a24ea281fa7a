#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "gemm_x3 or yolox" 2>&1 | tail -5
for k in 0 4 8 16; do echo "== GELU16_MAX_KITERS=$k"; GDRN_X3_GELU16_MAX_KITERS=$k timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "fc1" | cut -c1-120; done
echo "== trace k=4"; GDRN_GEMM_TRACE=1 timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "trace" | grep "epi=1" | sort | uniq -c | sort -rn | awk '{ if (!seen[$6$7$8]++) print }' | head -6 | cut -c1-420
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>&1 | tail -1 | cut -c1-900
