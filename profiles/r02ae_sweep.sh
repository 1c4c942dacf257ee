set -u
OUT=gpurun_out; mkdir -p $OUT
for FX in 4 8; do for SK in 0.25 0.30 0.35 0.40; do
  TMD_NL_FX=$FX TMD_BENCH_CONFIGS=0 timeout 200 python bench.py --workload lj1m --quick --steps 90 --skin $SK > $OUT/r02ae_fx${FX}_sk$SK.json 2> $OUT/r02ae.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02ae_fx${FX}_sk$SK.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("FX=$FX skin=$SK", "ms_per_step=%.4f"%d["ms_per_step"], "fused launch_ms=%.4f"%r["launch_ms"], d["config"]["nlist"]["builds"], "builds")
PY
done; done
