#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm_x3 or mlp_fused_x3" 2>&1 | tail -3
for c in 0 1 0 1; do echo "== COLLECT=$c"; GDRN_X3_COLLECT=$c timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "'shape'" | cut -c1-110; done
