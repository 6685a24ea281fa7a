#!/bin/bash
# compute-sanitizer memcheck over the kernels added / reworked this round (slow: a few tests only)
cd "$(dirname "$0")/.."
export GDRN_PDL=${GDRN_PDL:-1}
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "mlp_fused_x3 or dwconv_ln_pingpong or gemm_x3_vs_fp64" 2>&1 | tail -15
echo "exit=$?"
