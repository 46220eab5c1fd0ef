set -x
for f in 264 280 288 292 296; do
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f --lanes 512 2>/dev/null | tail -1 > gpurun_out/t.json
  python -c "import json;d=json.load(open('gpurun_out/t.json'));print('cfg5 rows frames $f',round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()})"
done
