# round 2, final build: full GPU suite, smoke, default bench line, reference arm
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py --impl reference --steps 10 --warmup 2 | tail -1 > gpurun_out/r02_bench_reference_cfg2.json
python bench.py --steps 20 --warmup 3 2>gpurun_out/r02_bench_default.err | tail -1 > gpurun_out/r02_bench_default.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_default.json'));r=json.load(open('gpurun_out/r02_bench_reference_cfg2.json'));print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'ref',round(r['value']),'e2e/ref',round(d['e2e']['value']/r['value'],2),'cpu_baseline',round(d['cpu_baseline']['value']),'launches',d['gpu_launches'],'clocks',d['clocks'],'streaming',d['streaming'],'strong',d['strong_cfg5']['ms_per_step'])"
