#!/bin/bash
# Every bench line of the round on one GPU (plain runs, no profiler):  gpurun -- bash profiles/final_bench.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
for W in lj1m lj32k lj256k dimer256k lj1m_langevin replicas; do
  timeout 400 python bench.py --workload $W > $OUT/final_${TAG}_$W.json 2> $OUT/final_${TAG}_$W.err
  tail -c 300 $OUT/final_${TAG}_$W.json | head -c 10 > /dev/null
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/final_${TAG}_$W.json").read().strip().splitlines()[-1])
    print("$W", "value=%.4g" % d["value"], d["unit"], "ms_per_step=%.4f" % d["ms_per_step"], "roofline.frac=%.3f" % d.get("roofline",{}).get("frac",float("nan")), "e2e=%.4g" % d.get("e2e",{}).get("value",float("nan")))
except Exception as e:
    print("$W FAILED", e)
PY
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/final_${TAG}_reference.json 2> $OUT/final_${TAG}_reference.err
tail -c 400 $OUT/final_${TAG}_reference.json
