python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 4096 8192; do for l in 0 1; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload cfg1 --frames $f --lanes $l 2>/dev/null | tail -1 > gpurun_out/t_c1_${f}_$l.json
  python -c "import json;d=json.load(open('gpurun_out/t_c1_${f}_$l.json'));print('cfg1 frames $f lanes $l',d['config'].get('kernel'),round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done; done
for l in 0 104; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload run_simple --frames 4096 --lanes $l 2>/dev/null | tail -1 > gpurun_out/t_rs_$l.json
  python -c "import json;d=json.load(open('gpurun_out/t_rs_$l.json'));print('run_simple 4096 frames lanes $l',d['config'].get('kernel'),round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done
