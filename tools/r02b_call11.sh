#!/bin/bash
# 2-row conv tile of the depthwise ping-pong kernel (GDRN_DW_R2=1) + 32-slice FC: bit-exactness, per-stage times, whole-step A/B
cd "$(dirname "$0")/.."
GDRN_DW_R2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "dwconv_ln_pingpong or north_star or b64 or tiny_small" 2>&1 | tail -3
echo "--- R2=0"; GDRN_DW_R2=0 timeout 300 python tools/dw_bench.py 2>&1 | grep "variant=1"
echo "--- R2=1"; GDRN_DW_R2=1 timeout 300 python tools/dw_bench.py 2>&1 | grep "variant=1"
STEPS=30 tools/ab.sh "GDRN_DW_R2=0" "GDRN_DW_R2=1" "GDRN_DW_R2=0" "GDRN_DW_R2=1"
