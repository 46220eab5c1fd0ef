set -x
for f in 136 148 160; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f --lanes 256 2>/dev/null | tail -1 > gpurun_out/t.json
  python -c "import json;d=json.load(open('gpurun_out/t.json'));print('cfg5 hist frames $f',round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()},d['config']['kernel'])"
done
for f in 880 888 896; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f --lanes 512 2>/dev/null | tail -1 > gpurun_out/t.json
  python -c "import json;d=json.load(open('gpurun_out/t.json'));print('cfg5 rows frames $f',round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()},d['config']['kernel'])"
done
