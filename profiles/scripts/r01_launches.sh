set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_cfg2.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2_default.json 2> gpurun_out/bench_cfg2_default.err
tail -c 2500 gpurun_out/bench_cfg2_default.json
