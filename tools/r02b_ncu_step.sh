#!/bin/bash
# One parity-mode step under ncu (run under gpurun): per-launch duration, tensor-pipe activity and DRAM bytes of every kernel of
# the forward -> gpurun_out/step_launches.csv; summarise with tools/step_tensor_share.py (time-weighted tensor-pipe % of the step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GDRN_BENCH_GRAPHS=0
# kernels of the forward (weight packing at load time excluded); LPS = launches per forward at B = 64 in parity mode
MODEL_KERNELS='regex:gemm_|mlp_fused|dwconv|gn_|upsample|ln_patchify|stem_patchify|cast_|pose_lift|fc_f32'
LPS=${LPS:-148}
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second \
    --clock-control none --kernel-name-base demangled -k "$MODEL_KERNELS" -s $((2 * LPS)) -c $LPS --csv --log-file gpurun_out/step_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_step.log 2>&1
tail -2 gpurun_out/b_ncu_step.log | cut -c1-300
python tools/step_tensor_share.py gpurun_out/step_launches.csv | tail -45
