set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02n_gputests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r02n_gputests.log
timeout 300 python bench.py --workload lj1m --quick > $OUT/r02n_bench_lj1m.json 2> $OUT/r02n_bench_lj1m.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n_bench_lj1m.json").read().strip().splitlines()[-1])
print("lj1m", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["launch_ms"], d["roofline"]["step_roofline_frac"])
PY
