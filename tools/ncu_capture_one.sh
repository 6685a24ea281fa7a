#!/bin/bash
# usage (under gpurun): tools/ncu_capture_one.sh '<demangled kernel regex>' <skip> <tag>
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c 1 \
    -o gpurun_out/prof_$3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
