#!/bin/bash
# compute-sanitizer memcheck over the kernels added in the second session of round 2 (k-split schedule, split-K FC stack,
# narrow tiles of the small-batch forward)
cd "$(dirname "$0")/.."
GDRN_X3_KSPLIT=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "(ksplit and 20000) or (ksplit and 4096) or (north_star_tolerance and 5) or (gemm_x3_vs_fp64 and 1280)" 2>&1 | tail -12
echo "exit=$?"
