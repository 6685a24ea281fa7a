#!/bin/bash
# round-2 baseline: launch list of one parity-mode (bf16x3) step + a short parity-mode bench, same box
cd "$(dirname "$0")/.."
export GDRN_BENCH_GRAPHS=0
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 900 --csv \
    --log-file gpurun_out/launches_x3.csv python bench.py --precision bf16x3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_x3.log 2>&1
unset GDRN_BENCH_GRAPHS
timeout 300 python bench.py --precision bf16x3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_x3_base.json 2> gpurun_out/bench_x3_base.err
tail -c 600 gpurun_out/bench_x3_base.json
