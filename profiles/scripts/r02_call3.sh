# round 2, GPU call 3: two-slot / two-stream batch calls: regression, pipelining tests, bench lines
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 --workload cfg2 2>gpurun_out/r02_bench_cfg2.err | tail -1 > gpurun_out/r02_bench_cfg2.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_cfg2.json'));print(json.dumps({k:d[k] for k in ['value','ms_per_step','ms_per_step_serial','stage_ms','e2e','e2e_roofline','roofline','strong_cfg5','cpu_baseline','gpu_launches']},indent=1))"
tail -5 gpurun_out/r02_bench_cfg2.err
python bench.py --impl reference --steps 5 --warmup 1 --workload cfg2 | tail -1 > gpurun_out/r02_bench_reference_cfg2.json
cat gpurun_out/r02_bench_reference_cfg2.json
WORKLOADS="cfg1 cfg3 cfg4" bash profiles/scripts/r02_quick.sh
