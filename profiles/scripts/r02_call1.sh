# round 2, GPU call 1: micro-benchmarks, L2 fetch granularity on the traceback, ncu --set full of the config-3 default kernel
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
./profiles/microbench/hist_mix2 > gpurun_out/r02_hist_mix2.txt 2>&1
cat gpurun_out/r02_hist_mix2.txt
for g in 64 32; do
  VITB_L2_FETCH=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg2 2>/dev/null | tail -1 > gpurun_out/r02_cfg2_l2fetch$g.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg2_l2fetch$g.json'));print('cfg2 l2fetch $g',round(d['value']),round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['stage_ms'].items()},round(d['e2e']['value']))"
  VITB_L2_FETCH=$g python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg3 2>/dev/null | tail -1 > gpurun_out/r02_cfg3_l2fetch$g.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg3_l2fetch$g.json'));print('cfg3 l2fetch $g',round(d['value']),round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['stage_ms'].items()},round(d['e2e']['value']))"
done
ncu --set full --clock-control none --import-source on -k regex:acs_hist_group -s 3 -c 1 -f -o gpurun_out/r02_prof_cfg3_histgroup python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/r02_ncu_cfg3.log 2>&1
tail -3 gpurun_out/r02_ncu_cfg3.log
ls -la gpurun_out/*.ncu-rep | tail -3
