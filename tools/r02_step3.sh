#!/bin/bash
# round-2 step 3: new PnP / predictor tests + per-kernel launch list of one parity-mode step
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pnp or predictor" 2>&1 | tail -25 > gpurun_out/step3_tests.log
tail -12 gpurun_out/step3_tests.log
export GDRN_BENCH_GRAPHS=0
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:gemm_|mlp_fused|dwconv|gn_|upsample|ln_patchify|stem_patchify|cast_|pose_lift|fc_f32" -s 480 -c 170 --csv \
    --log-file gpurun_out/launches_x3_step3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_x3_step3.log 2>&1
tail -2 gpurun_out/b_ncu_x3_step3.log | cut -c1-300
