#!/bin/bash
# usage: tools/ncu_capture.sh  (run under gpurun) -- ncu --set full captures of the hottest kernels of one bench step
cd "$(dirname "$0")/.."
cap() {  # name-regex skip tag
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c 1 \
      -o gpurun_out/prof_$3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$3.log 2>&1
  tail -1 gpurun_out/ncu_$3.log
}
cap 'dwconv_ln_cluster_kernel<\(int\)16' 110 dw
cap 'gemm_tc_kernel<\(int\)256, \(int\)2>' 110 fc2
cap 'gemm_tc_kernel<\(int\)256, \(int\)1>' 122 fc1
cap 'gemm_tc_kernel<\(int\)256, \(int\)3>' 38 conv64
cap 'gn_gelu_kernel' 38 gn
