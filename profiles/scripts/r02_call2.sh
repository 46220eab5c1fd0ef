# round 2, GPU call 2 (the container of call 1 was lost with its gpurun_out): regression, micro-benchmark, L2 fetch granularity, bench lines
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
./profiles/microbench/hist_mix2 > gpurun_out/r02_hist_mix2.txt 2>&1
cat gpurun_out/r02_hist_mix2.txt
for g in 64 32; do
  VITB_L2_FETCH=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg2 2>/dev/null | tail -1 > gpurun_out/r02_cfg2_l2fetch$g.json
  python -c "import json;d=json.load(open('gpurun_out/r02_cfg2_l2fetch$g.json'));print('cfg2 l2fetch $g',round(d['value']),round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['stage_ms'].items()},round(d['e2e']['value']))"
done
WORKLOADS="cfg1 cfg3 cfg4 cfg5 run_simple" bash profiles/scripts/r02_quick.sh
