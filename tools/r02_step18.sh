#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tiny_small" 2>&1 | tail -25 | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "forward_vs_oracle or dwconv or gemm or north_star" 2>&1 | tail -5
