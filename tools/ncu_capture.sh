#!/bin/bash
# usage: tools/ncu_capture.sh  (run under gpurun) -- launch list + DRAM traffic of one bench step and ncu --set full
# captures of the hottest kernels.  Outputs under gpurun_out/; summarise with tools/summarize_ncu.py / make_traffic.py.
cd "$(dirname "$0")/.."
export GDRN_BENCH_GRAPHS=0
MODEL_KERNELS='regex:gemm_|mlp_fused|dwconv|gn_|upsample|ln_patchify|stem_patchify|cast_|pose_lift|fc_f32'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --kernel-name-base demangled -k "$MODEL_KERNELS" -c 480 --csv --log-file gpurun_out/traffic.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
cp gpurun_out/traffic.csv gpurun_out/launches.csv
cap() {  # name-regex skip tag
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c 1 \
      -o gpurun_out/prof_$3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$3.log 2>&1
  tail -1 gpurun_out/ncu_$3.log | cut -c1-120
}
cap 'dwconv_ln_cluster_kernel<\(int\)16' 110 dw
cap 'gemm_pair_kernel<\(int\)2, \(int\)8>' 90 fc2_pair
cap 'gemm_pair_kernel<\(int\)1, \(int\)16>' 110 fc1_pair16
cap 'gemm_tc_kernel<\(int\)256, \(int\)3>' 41 conv64
cap 'mlp_fused_kernel' 10 mlp_fused
