#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "online_targets or yolox" 2>&1 | tail -15
echo "== traces"; GDRN_GEMM_TRACE=1 timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "trace" | grep "epi=1" | sort | uniq -c | sort -rn | awk '{ if (!seen[$6$7$8]++) print }' | head -6 | cut -c1-420
