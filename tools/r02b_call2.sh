#!/bin/bash
# k-split schedule of the pair-x3 residual GEMMs: tests with the schedule on, then same-box A/B of the whole step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GDRN_X3_KSPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "ksplit or gemm_x3 or b64 or deterministic or sharded or north_star" 2>&1 | tail -5
GDRN_X3_KSPLIT=1 timeout 300 python tools/x3_shapes_bench.py 2>&1 | tail -14
echo "--- whole tiles"
GDRN_X3_KSPLIT=0 timeout 300 python tools/x3_shapes_bench.py 2>&1 | tail -14
STEPS=30 tools/ab.sh "GDRN_X3_KSPLIT=0" "GDRN_X3_KSPLIT=1" "GDRN_X3_KSPLIT=0" "GDRN_X3_KSPLIT=1"
