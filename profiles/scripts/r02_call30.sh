python -m pytest tests/test_gpu_history_k9.py -x -q 2>&1 | tail -3
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_history_k9.py 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -4
for l in 1 104; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload run_simple --lanes $l 2>/dev/null | tail -1 > gpurun_out/t_rs_$l.json
  python -c "import json;d=json.load(open('gpurun_out/t_rs_$l.json'));print('run_simple lanes $l',d['config'].get('kernel'),round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done
for f in 4096 8192 12288 16384; do for l in 1 104; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload cfg1 --frames $f --lanes $l 2>/dev/null | tail -1 > gpurun_out/t_c1_${f}_$l.json
  python -c "import json;d=json.load(open('gpurun_out/t_c1_${f}_$l.json'));print('cfg1 frames $f lanes $l',d['config'].get('kernel'),round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done; done
