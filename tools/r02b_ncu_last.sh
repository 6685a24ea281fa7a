#!/bin/bash
# ncu --set full of the kernels reworked at the end of the session: depthwise ping-pong kernel with the two-row thread tile,
# chamfer forward with four queries per thread
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:dwconv_ln_pp_kernel<\(int\)8, \(bool\)1' -s 2 -c 1 -o gpurun_out/prof_dw_pp_r2 -f python tools/dw_bench.py > gpurun_out/ncu_dw_r2.log 2>&1
tail -1 gpurun_out/ncu_dw_r2.log | cut -c1-100
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:nnd_fwd' -s 3 -c 1 \
   -o gpurun_out/prof_nnd_q4 -f python bench.py --workload nnd > gpurun_out/ncu_nnd_q4.log 2>&1
tail -1 gpurun_out/ncu_nnd_q4.log | cut -c1-100
