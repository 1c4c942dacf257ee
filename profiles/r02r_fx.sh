set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02r_gputests.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02r_gputests.log
for FX in 1 2 4 8; do
  TMD_NL_FX=$FX timeout 300 python bench.py --workload lj1m --quick --steps 60 > $OUT/r02r_bench_fx$FX.json 2> $OUT/r02r_bench_fx$FX.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02r_bench_fx$FX.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("FX=$FX", "ms_per_step=%.4f"%d["ms_per_step"], "fused launch_ms=%.4f"%r["launch_ms"], "frac=%.4f"%r["frac"], "eval_ms=%.4f"%r["force_evaluation"]["launch_ms"], d["config"]["nlist"])
PY
done
CMD="python bench.py --workload lj1m --steps 30 --warmup 3 --quick"
timeout 200 $CMD > $OUT/plain_lj1m.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02r_launches_lj1m.csv $CMD > $OUT/ncu_launches_lj1m.log 2>&1
echo "ncu rc=$?"
