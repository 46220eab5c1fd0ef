# regenerates the round-1 numbers under gpurun_out/ (copied into profiles/ by hand): every BASELINE.json config, launch list, full captures
set -x
for w in cfg1 cfg2 cfg3 cfg4 cfg5 run_simple; do python bench.py --steps 10 --warmup 3 --workload $w 2>/dev/null | tail -1 > gpurun_out/bench_$w.json; done
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_cfg2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_hist -s 3 -c 1 -f -o gpurun_out/prof_cfg2_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg2_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceback_hist_seg -s 3 -c 1 -f -o gpurun_out/prof_cfg2_tb_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg2_tb_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_hist -s 3 -c 1 -f -o gpurun_out/prof_cfg1_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg1 > gpurun_out/ncu_cfg1_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_group -s 3 -c 1 -f -o gpurun_out/prof_cfg3_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/ncu_cfg3_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:acs_cta -s 3 -c 1 -f -o gpurun_out/prof_cfg5_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/ncu_cfg5_final.log 2>&1
ls -la gpurun_out/*.ncu-rep
