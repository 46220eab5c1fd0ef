# round 2, GPU call 5: K=15 history kernel with 4-byte table entries; 1 vs 2 CTAs per SM; against the decision-row kernel
set -x
python -m pytest tests/test_gpu_history_k15.py tests/test_gpu_pipelining.py -x -q 2>&1 | tail -3
for m in 1 2; do
  VITB_HC_MINB=$m python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 2>/dev/null | tail -1 > gpurun_out/r02_cfg5_minb$m.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg5_minb$m.json'));print('cfg5 hist minb $m',round(d['value']),round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()},d['config']['kernel'])"
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --lanes 512 2>/dev/null | tail -1 > gpurun_out/r02_cfg5_rows.json
python -c "import json;d=json.load(open('gpurun_out/r02_cfg5_rows.json'));print('cfg5 rows',round(d['value']),round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()},d['config']['kernel'])"
for f in 128 256; do
 for m in 1 2; do
  VITB_HC_MINB=$m python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f 2>/dev/null | tail -1 > gpurun_out/r02_cfg5_f${f}_minb$m.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg5_f${f}_minb$m.json'));print('cfg5 frames $f hist minb $m',round(d['value']),round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()})"
 done
 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --workload cfg5 --frames $f --lanes 512 2>/dev/null | tail -1 > gpurun_out/r02_cfg5_f${f}_rows.json
 python -c "import json;d=json.load(open('gpurun_out/r02_cfg5_f${f}_rows.json'));print('cfg5 frames $f rows',round(d['value']),round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['stage_ms'].items()})"
done
