set -x
for c in 4 8 16; do
  VITB_E2E_CHUNKS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload cfg2 2>/dev/null | tail -1 > gpurun_out/t.json
  python -c "import json;d=json.load(open('gpurun_out/t.json'));print('cfg2 e2e chunks $c',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),round(d['e2e_roofline']['copy_ms_per_step'],3))"
done
for c in 4 8; do
  VITB_E2E_CHUNKS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --workload cfg3 2>/dev/null | tail -1 > gpurun_out/t.json
  python -c "import json;d=json.load(open('gpurun_out/t.json'));print('cfg3 e2e chunks $c',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),round(d['e2e_roofline']['copy_ms_per_step'],3))"
done
