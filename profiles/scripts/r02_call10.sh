set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_history_k15.py -x -q -k "Cassini or k15 or cfg5 or streaming or K15" 2>&1 | tail -4
for f in 1024 888 264; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f 2>/dev/null | tail -1 > gpurun_out/r02_cfg5_f${f}.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg5_f${f}.json'));print('cfg5 frames $f',round(d['value']),round(d['ms_per_step'],3),round(d['ms_per_step_serial'],3),{k:round(v,3) for k,v in d['stage_ms'].items()},d['config']['kernel'])"
done
