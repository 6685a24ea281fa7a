#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fps" 2>&1 | tail -3
for m in 0 16384; do echo "CLUSTER_MIN_PN=$m"; GDRN_FPS_CLUSTER_MIN_PN=$m timeout 300 python bench.py --workload fps 2>&1 | grep '"op"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['ms'],4), 'ms')"; done
bash tools/r02_sanitize.sh
