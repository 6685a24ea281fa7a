#!/bin/bash
cd "$(dirname "$0")/.."
echo "== rational erf"; timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "'shape'" 
echo "== libm erff"; GDRN_X3_GELU_MODE=4 timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "fc1"
echo "== traces"; GDRN_GEMM_TRACE=1 timeout 300 python tools/x3_shapes_bench.py 2>&1 | grep "trace" | sort | uniq -c | sort -rn | head -30
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "renderer_surfaces or multi_mesh or depth_refine or rasteriser" 2>&1 | tail -5
