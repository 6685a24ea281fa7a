#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ransac_voting_layer or estimate_voting or voting" 2>&1 | tail -15
