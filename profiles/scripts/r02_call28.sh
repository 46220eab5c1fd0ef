python -m pytest tests/test_gpu_parity.py tests/test_gpu_tag_stress.py tests/test_gpu_simd_sat.py tests/test_gpu_reference_programs.py tests/test_gpu_examples.py tests/test_gpu_generic.py tests/test_gpu_vs_reference.py -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong 2>/dev/null | tail -1 > gpurun_out/t_cfg2.json
python -c "import json;d=json.load(open('gpurun_out/t_cfg2.json'));print('mapped',d['streaming'])"
VITB_NO_MAPPED_STREAMING=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong 2>/dev/null | tail -1 > gpurun_out/t_cfg2b.json
python -c "import json;d=json.load(open('gpurun_out/t_cfg2b.json'));print('copies',d['streaming'])"
