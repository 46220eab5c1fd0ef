# bench lines of every BASELINE.json config (+ reference arm, launch list) without the ncu --set full captures of r01_all.sh
set -x
for w in cfg1 cfg2 cfg3 cfg4 cfg5 run_simple; do python bench.py --steps 10 --warmup 3 --workload $w 2>/dev/null | tail -1 > gpurun_out/bench_$w.json; done
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_cfg2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_cfg2.log 2>&1
