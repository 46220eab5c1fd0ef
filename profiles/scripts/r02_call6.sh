# round 2, GPU call 6: K=15 history kernel with the tag added once per butterfly on the other pipe
set -x
python -m pytest tests/test_gpu_history_k15.py tests/test_gpu_pipelining.py -x -q 2>&1 | tail -15
for f in 1024 128; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f 2>/dev/null | tail -1 > gpurun_out/r02_cfg5_f${f}_hist.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg5_f${f}_hist.json'));print('cfg5 frames $f hist',round(d['value']),round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()})"
done
