#!/bin/bash
# round-2 ncu evidence (run under gpurun): launch list + DRAM traffic of one parity-mode step, ncu --set full captures of the
# hot kernels of the dense path and of every native op north_star names.  Outputs under gpurun_out/; summarise with
# tools/summarize_ncu.py r02 and tools/make_traffic.py r02_traffic_x3.json traffic_x3.csv <git sha>.
cd "$(dirname "$0")/.."
export GDRN_BENCH_GRAPHS=0
MODEL_KERNELS='regex:gemm_|mlp_fused|dwconv|gn_|upsample|ln_patchify|stem_patchify|cast_|pose_lift|fc_f32'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --kernel-name-base demangled -k "$MODEL_KERNELS" -s 340 -c 360 --csv --log-file gpurun_out/traffic_x3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_x3.log 2>&1
cp gpurun_out/traffic_x3.csv gpurun_out/launches.csv
cap() {  # name-regex skip tag  (dense path: one launch of the named kernel out of a bench run)
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c 1 \
      -o gpurun_out/prof_$3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$3.log 2>&1
  tail -1 gpurun_out/ncu_$3.log | cut -c1-100
}
cap 'gemm_pair_x3_kernel<\(int\)256, \(int\)2,' 50 x3_fc2_s2
cap 'gemm_pair_x3_kernel<\(int\)256, \(int\)1, \(int\)8>' 40 x3_fc1_s2
cap 'gemm_pair_x3_kernel<\(int\)256, \(int\)1, \(int\)16>' 4 x3_fc1_s1
cap 'mlp_fused_x3_kernel' 4 x3_mlp_fused_s0
cap 'gemm_pair_x3_kernel<\(int\)256, \(int\)3,' 16 x3_conv64
cap 'dwconv_ln_pp_kernel' 50 dw_pp
capn() {  # native ops: kernel-regex workload tag
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s 3 -c 1 \
      -o gpurun_out/prof_$3 -f python bench.py --workload $2 > gpurun_out/ncu_$3.log 2>&1
  tail -1 gpurun_out/ncu_$3.log | cut -c1-100
}
capn 'fps_kernel' fps fps
capn 'vote_kernel<\(bool\)0, \(bool\)0>' voting vote_mask
capn 'vote_kernel<\(bool\)0, \(bool\)1>' voting vote_count
capn 'nnd_fwd' nnd nnd
capn 'flow_kernel' flow flow
capn 'rast_tri' raster rast_tri
capn 'depth_refine_kernel' raster refine
capn 'upnp_kernel' upnp upnp
