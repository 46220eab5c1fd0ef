ncu --set full --clock-control none --import-source on -k regex:acs_hist_cta -c 1 -o gpurun_out/r02_hist_cta_v3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --workload cfg5 --frames 148 --lanes 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_cta_kernel -c 1 -o gpurun_out/r02_cta_v3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --workload cfg5 --frames 296 --lanes 512 > /dev/null 2>&1
ncu -i gpurun_out/r02_hist_cta_v3.ncu-rep --page details > gpurun_out/r02_ncu_full_acs_hist_cta_cfg5_148frames.txt 2>&1
ncu -i gpurun_out/r02_cta_v3.ncu-rep --page details > gpurun_out/r02_ncu_full_acs_cta_cfg5_296frames.txt 2>&1
grep -i "^    Duration\|Issue Slots Busy\|dram__bytes\|Registers Per" gpurun_out/r02_ncu_full_acs_hist_cta_cfg5_148frames.txt gpurun_out/r02_ncu_full_acs_cta_cfg5_296frames.txt
