#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 8 16 24 32 40 56; do echo "== DBG=$d"; GDRN_DW_DBG=$d timeout 200 python tools/dw_bench.py 2>&1 | grep "variant=1" | grep "split=1"; done
