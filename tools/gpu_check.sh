#!/bin/bash
# tools/gpu_check.sh -- what every GPU visit runs: parity tests, then the bench (CPU baseline only when FULL=1).
# VARIANTS="A=1 B=2+C=3" reruns the bench once per extra environment setting (tuning knobs; + joins several).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"batch",round(d["e2e_batch_api"]["value"]),'ms/step',round(d['ms_per_step'],3),'kernels',{k:round(v,3) for k,v in d['roofline']['all_kernels_ms'].items()},'clk',d['clocks']['sm_mhz'],d['clocks']['reasons'])
PY
}
if [ "$FULL" = "1" ]; then python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; else python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; fi
echo bench_exit=$?; tail -3 gpurun_out/bench.err; summ gpurun_out/bench.json
i=0
for v in $VARIANTS; do
  i=$((i+1)); env ${v//+/ } python bench.py --no-cpu-baseline > gpurun_out/bench_var$i.json 2> gpurun_out/bench_var$i.err; echo "variant $v exit=$?"; summ gpurun_out/bench_var$i.json
done
