set -x
python -m pytest tests/test_gpu_history_k15.py tests/test_gpu_tag_stress.py -x -q -k "k15 or Cassini" 2>&1 | tail -3
for f in 148 1024; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f 2>/dev/null | tail -1 > gpurun_out/t.json
  python -c "import json;d=json.load(open('gpurun_out/t.json'));print('cfg5 frames $f',round(d['value']),round(d['ms_per_step_serial'],3),{k:round(v,3) for k,v in d['stage_ms'].items()})"
done
