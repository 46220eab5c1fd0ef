set -x
ncu --set full --clock-control none --import-source on -k regex:acs_cta -s 1 -c 1 -f -o gpurun_out/prof_cfg5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg5 --frames 296 --lanes 512 > gpurun_out/ncu_cfg5.log 2>&1
tail -2 gpurun_out/ncu_cfg5.log
