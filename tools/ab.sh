#!/bin/bash
# usage (under gpurun): tools/ab.sh "ENV1=.. ENV2=.." "ENV.." ...  -> one short bench line per environment, same box
cd "$(dirname "$0")/.."
for envs in "$@"; do
  env $envs timeout 200 python bench.py --no-cpu-baseline --steps ${STEPS:-30} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-44s value %8.1f  ms %.3f  gemm %.3f dw %.3f other %.3f  x3 %s' % ('$envs', d['value'], d['ms_per_step'], r['gemm_ms_per_step'], r['dwconv_ms_per_step'], r['other_ms_per_step'], (d.get('fp32_parity_mode') or {}).get('value')))"
done
