for l in 0 1 2 4; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload run_simple --lanes $l 2>/dev/null | tail -1 > gpurun_out/t_rs_$l.json
  python -c "import json;d=json.load(open('gpurun_out/t_rs_$l.json'));print('run_simple lanes $l',d['config'].get('kernel'),round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done
for f in 1024 2048 4096; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload cfg2 --frames $f 2>/dev/null | tail -1 > gpurun_out/t_c2_$f.json
  python -c "import json;d=json.load(open('gpurun_out/t_c2_$f.json'));print('cfg2 frames $f',d['config'].get('kernel'),round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done
