#!/bin/bash
# small-batch forward (B = 5): graph-replay time and one-forward ncu launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/fwd_small.py 5 3 2>&1 | tail -2
timeout 300 python tools/fwd_small.py 1 3 2>&1 | tail -1
timeout 300 python tools/fwd_small.py 16 3 2>&1 | tail -1
MODEL_KERNELS='regex:gemm_|mlp_fused|dwconv|gn_|upsample|ln_patchify|stem_patchify|cast_|pose_lift|fc_f32'
FWD_SMALL_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --kernel-name-base demangled -k "$MODEL_KERNELS" -s 400 -c 200 --csv --log-file gpurun_out/small_launches.csv \
    python tools/fwd_small.py 5 4 > gpurun_out/ncu_small.log 2>&1
python tools/step_tensor_share.py gpurun_out/small_launches.csv | head -45
