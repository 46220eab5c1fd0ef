set -x
python -m pytest tests/test_gpu_history.py tests/test_gpu_tag_stress.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for w in cfg2 cfg1 cfg4 run_simple; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong --workload $w 2>/dev/null | tail -1 > gpurun_out/t_$w.json
  python -c "import json;d=json.load(open('gpurun_out/t_$w.json'));print('$w',round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done
