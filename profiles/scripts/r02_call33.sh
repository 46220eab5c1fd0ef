timeout 70 python -m pytest tests/test_gpu_history_k9.py -x -q -k "IS-95A or Voyager or rate_one_third" 2>&1 | tail -2
timeout 45 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg3 2>/dev/null | tail -1 > gpurun_out/t_cfg3.json
python -c "import json;d=json.load(open('gpurun_out/t_cfg3.json'));print('cfg3',d['config'].get('kernel'),round(d['value']),round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
