set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q > $OUT/r02z_gputests.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02z_gputests.log
timeout 600 python bench.py > $OUT/r02z_bench_default.json 2> $OUT/r02z_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z_bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("lj1m", d["value"], d["ms_per_step"], r["kernel"], r["frac"], r["launch_ms"], r["step_roofline_frac"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"], d["e2e"]["value"], d["gpu_launches"])
for c in d.get("configs", []): print(c.get("workload"), c.get("value"), c.get("ms_per_step"), (c.get("roofline") or {}).get("frac"), c.get("error"))
PY
