#!/bin/bash
# tools/ab.sh "ENV_A" "ENV_B" ... -- alternate device-only bench runs of several environment settings (A/B timing on one box)
mkdir -p gpurun_out
for round in 1 2; do
  i=0
  for v in "$@"; do
    i=$((i+1))
    env ${v//+/ } python bench.py --no-cpu-baseline --files 0 --steps 20 > gpurun_out/ab_${i}_${round}.json 2> gpurun_out/ab_${i}_${round}.err
    python - gpurun_out/ab_${i}_${round}.json "$v" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('%-40s ms/step %.3f' % (sys.argv[2], d['ms_per_step']), {k.split('(')[0][:8]:round(v,3) for k,v in d['roofline']['all_kernels_ms'].items()})
PY
  done
done
