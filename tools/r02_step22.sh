#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "dwconv or tiny_small or north_star" 2>&1 | tail -4
timeout 200 python tools/dw_bench.py 2>&1 | grep "variant=1"
GDRN_DW_TRACE=1 timeout 200 python tools/dw_bench.py 2>&1 | grep "pp trace" | grep "split=1" | sort | uniq -c | sort -rn | awk '{ if (!seen[$5$6]++) print }' | cut -c1-330
