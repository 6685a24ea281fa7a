#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "online_targets" 2>&1 | tail -12
bash tools/r02_ncu_capture.sh 2>&1 | tail -20
ls -la gpurun_out/*.ncu-rep | wc -l
