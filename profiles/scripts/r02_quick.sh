# quick regression + per-config bench lines (no CPU baseline) into gpurun_out/r02_quick_*.json
#python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in ${WORKLOADS:-cfg1 cfg2 cfg3 cfg4 cfg5 run_simple}; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload $w 2>/dev/null | tail -1 > gpurun_out/r02_quick_$w.json
  python -c "import json,sys;d=json.load(open('gpurun_out/r02_quick_$w.json'));print('$w',round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()},round(d['e2e']['value']),d['config']['kernel'])"
done
