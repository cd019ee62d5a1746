mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1)); env ${v//+/ } python bench.py --no-cpu-baseline --files 0 --steps 10 > gpurun_out/var$i.json 2> gpurun_out/var$i.err; echo "variant $v exit=$?"
  python - gpurun_out/var$i.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'batch',round(d['e2e_batch_api']['value']),'ms/step',round(d['ms_per_step'],3),{k.split('(')[0][:8]:round(v,3) for k,v in d['roofline']['all_kernels_ms'].items()})
PY
done
