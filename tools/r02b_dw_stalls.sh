#!/bin/bash
# source-level stall profile of the stage-2 depthwise ping-pong kernel (split output)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:dwconv_ln_pp_kernel<\(int\)8, \(bool\)1' -s 2 -c 1 -o gpurun_out/prof_dw8 -f python tools/dw_bench.py > gpurun_out/ncu_dw8.log 2>&1
tail -3 gpurun_out/ncu_dw8.log | cut -c1-200
python tools/ncu_stalls.py gpurun_out/prof_dw8.ncu-rep 70
