# round 2: bench lines of every configuration (1 GPU) into gpurun_out/r02_bench_*.json, reference arm of config 2
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_bench_default.json
python bench.py --impl reference --steps 10 --warmup 2 | tail -1 > gpurun_out/r02_bench_reference_cfg2.json
for w in cfg1 cfg3 cfg4 cfg5 run_simple; do
  python bench.py --steps 10 --warmup 3 --no-strong --workload $w 2>/dev/null | tail -1 > gpurun_out/r02_bench_$w.json
done
python bench.py --steps 3 --warmup 3 --no-strong --no-cpu-baseline --workload cfg5 --window 400 2>/dev/null | tail -1 > gpurun_out/r02_bench_cfg5_window400.json
python - <<'PY'
import json
for w in ["default", "cfg1", "cfg3", "cfg4", "cfg5", "run_simple", "cfg5_window400"]:
    d = json.load(open(f"gpurun_out/r02_bench_{w}.json"))
    cb = d.get("cpu_baseline", {}).get("value")
    print(w, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "serial", round(d["ms_per_step_serial"], 4),
          {k: round(v, 4) for k, v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["value"]), "e2e_frac", round(d["e2e_roofline"]["frac"], 3),
          "issue_frac", round(d["roofline"]["frac"], 3), "hbm_frac", round(d["roofline_hbm"]["frac"], 3), "cpu", cb and round(cb), d["config"]["kernel"],
          "ws", d["config"].get("workspace_bytes"))
d = json.load(open("gpurun_out/r02_bench_default.json"))
print(json.dumps(d.get("strong_cfg5")))
print(open("gpurun_out/r02_bench_reference_cfg2.json").read()[:400])
PY
