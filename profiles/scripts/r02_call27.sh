python -m pytest tests/test_gpu_history_k15.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tag_stress.py -x -q -k "Cassini or cassini or K15 or k15" 2>&1 | tail -3
for f in 128 148; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f --lanes 256 2>/dev/null | tail -1 > gpurun_out/t_cfg5_hist_$f.json
  python -c "import json;d=json.load(open('gpurun_out/t_cfg5_hist_$f.json'));print('hist $f',round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done
