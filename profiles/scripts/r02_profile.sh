# round 2 profiling pass (1 GPU): launch list of the default bench command, ncu --set full of the dominant kernels
set -x
mkdir -p gpurun_out/prof
# every launch of the default command with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/r02_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/prof/r02_launches_cfg2.log 2>&1
# config 2 ACS kernel (survivor-history, two frames per lane)
ncu --set full --clock-control none --import-source on -k regex:acs_hist_kernel -s 4 -c 1 -f -o gpurun_out/prof/r02_prof_cfg2_hist \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-pipelining > gpurun_out/prof/r02_ncu_cfg2.log 2>&1
# config 2 traceback (single chain per frame = the overlapped form) and segmented form
ncu --set full --clock-control none -k regex:traceback_hist -s 4 -c 2 -f -o gpurun_out/prof/r02_prof_cfg2_tb \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-pipelining > gpurun_out/prof/r02_ncu_cfg2_tb.log 2>&1
# config 3 ACS kernel (K = 9, frame over four lanes)
ncu --set full --clock-control none --import-source on -k regex:acs_hist_group -s 3 -c 1 -f -o gpurun_out/prof/r02_prof_cfg3_histgroup \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-pipelining --workload cfg3 > gpurun_out/prof/r02_ncu_cfg3.log 2>&1
# config 5: decision-row kernel (one wave of frame pairs) and frame-per-CTA history kernel (one wave of frames)
ncu --set full --clock-control none --import-source on -k regex:acs_cta_kernel -s 3 -c 1 -f -o gpurun_out/prof/r02_prof_cfg5_rows \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-pipelining --workload cfg5 --frames 296 --lanes 512 > gpurun_out/prof/r02_ncu_cfg5_rows.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_hist_cta -s 3 -c 1 -f -o gpurun_out/prof/r02_prof_cfg5_histcta \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-pipelining --workload cfg5 --frames 148 > gpurun_out/prof/r02_ncu_cfg5_hist.log 2>&1
ls -la gpurun_out/prof/
