# round 2, GPU call 7: full regression after the slot / K=15 selection changes; config 5 with the wave split
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for f in 1024 512 256 128; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f 2>/dev/null | tail -1 > gpurun_out/r02_cfg5_f${f}.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg5_f${f}.json'));print('cfg5 frames $f',round(d['value']),round(d['ms_per_step'],3),round(d['ms_per_step_serial'],3),{k:round(v,3) for k,v in d['stage_ms'].items()},d['config']['kernel'])"
done
