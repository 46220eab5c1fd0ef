# round 2, 2-GPU call: real multi-device tests, torchrun bench at N=2 with the strong config-5 leg and the in-process multi leg
set -x
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_reference_programs.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r02_bench_2gpu.err | tail -1 > gpurun_out/r02_bench_cfg2_2gpu.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_cfg2_2gpu.json'));print(json.dumps({k:d.get(k) for k in ['value','ms_per_step','ms_per_step_serial','e2e','e2e_roofline','strong_cfg5','multi_in_process']},indent=1))"
tail -3 gpurun_out/r02_bench_2gpu.err
