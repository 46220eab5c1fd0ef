set -x
./build/tmp/lds128_merge | tee gpurun_out/r02_lds128_merge.txt
ncu --set full --clock-control none --import-source on -k regex:acs_hist_cta -c 1 -o gpurun_out/r02_hist_cta_v2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --workload cfg5 --frames 148 --lanes 256 > /dev/null 2>&1
ncu -i gpurun_out/r02_hist_cta_v2.ncu-rep --page details > gpurun_out/r02_ncu_full_acs_hist_cta_cfg5_148frames_v2.txt 2>&1
grep -i "duration\|issue slots\|wavefront\|bank\|Mem Pipes\|registers per" gpurun_out/r02_ncu_full_acs_hist_cta_cfg5_148frames_v2.txt | head -30
