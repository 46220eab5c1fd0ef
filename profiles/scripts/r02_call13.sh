set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in cfg2 cfg4 cfg5; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload $w 2>gpurun_out/t.err | tail -1 > gpurun_out/t_$w.json || tail -5 gpurun_out/t.err
  python -c "import json;d=json.load(open('gpurun_out/t_$w.json'));print('$w',round(d['value']),round(d['ms_per_step'],4),d.get('streaming'))"
done
