# round 2, GPU call 4: what limits the overlap of traceback (batch i) and ACS (batch i+1)?  gentler traceback (fewer segments in flight)
set -x
python -m pytest tests/test_gpu_pipelining.py -x -q 2>&1 | tail -3
for t in 131072 65536 32768; do
  VITB_SEG_TARGET=$t python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong --workload cfg2 2>/dev/null | tail -1 > gpurun_out/r02_cfg2_seg$t.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg2_seg$t.json'));print('cfg2 seg_target $t',round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
done
VITB_NO_SEG_TRACEBACK=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong --workload cfg2 2>/dev/null | tail -1 > gpurun_out/r02_cfg2_noseg.json
python -c "import json;d=json.load(open('gpurun_out/r02_cfg2_noseg.json'));print('cfg2 noseg',round(d['value']),round(d['ms_per_step'],4),round(d['ms_per_step_serial'],4),{k:round(v,4) for k,v in d['stage_ms'].items()})"
WORKLOADS="cfg1 cfg3 cfg4 run_simple" bash profiles/scripts/r02_quick.sh
