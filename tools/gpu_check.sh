#!/bin/bash
# tools/gpu_check.sh -- what every GPU visit runs: parity tests, then the bench (no CPU baseline unless FULL=1)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_gpu.log
if [ "$FULL" = "1" ]; then python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; else python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; fi
echo bench_exit=$?; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'kernels',{k:round(v,3) for k,v in d['roofline']['all_kernels_ms'].items()},'clk',d['clocks'])
PY
