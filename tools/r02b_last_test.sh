#!/bin/bash
cd "$(dirname "$0")/.."
timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "uncertainty_pnp_v2 or upnp_vs_vendored" 2>&1 | tail -6
