set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02b_gputests.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/r02b_gputests.log
timeout 600 python bench.py > $OUT/r02b_bench_default.json 2> $OUT/r02b_bench_default.err; echo "bench rc=$?"
tail -c 600 $OUT/r02b_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench_default.json").read().strip().splitlines()[-1])
print("lj1m", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_roofline_frac"], d["clocks"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"])
for c in d.get("configs", []): print(c.get("workload"), c.get("value"), c.get("ms_per_step"), (c.get("roofline") or {}).get("frac"), c.get("error"), (c.get("clocks") or {}).get("samples"))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02b_bench_reference.json 2>&1; tail -c 400 $OUT/r02b_bench_reference.json
