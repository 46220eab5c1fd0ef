set -x
ncu --set full --clock-control none --import-source on -k regex:acs_ -s 3 -c 1 -f -o gpurun_out/prof_cfg2_v2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg2_v2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_ -s 3 -c 1 -f -o gpurun_out/prof_cfg3_v2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/ncu_cfg3_v2.log 2>&1
ls -la gpurun_out/*.ncu-rep
