# round 2: N-GPU run of the default bench (weak-scaling config 2 + strong-scaling config 5 + in-process multi-device leg)
set -x
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv | head -10
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r02_bench_${N}gpu.err | tail -1 > gpurun_out/r02_bench_cfg2_${N}gpu.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_cfg2_${N}gpu.json'));print(json.dumps({k:d.get(k) for k in ['n_gpus','value','ms_per_step','ms_per_step_serial','e2e','e2e_roofline','strong_cfg5','multi_in_process','clocks']},indent=1))"
tail -3 gpurun_out/r02_bench_${N}gpu.err
python -m pytest tests/test_gpu_multi_device.py -x -q 2>&1 | tail -2
