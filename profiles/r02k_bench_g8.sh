set -u
OUT=gpurun_out; mkdir -p $OUT
run() { # N workload tag
  N=$1; W=$2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --workload $W --steps ${3:-30} > $OUT/r02k_bench_${W}_g$N.json 2> $OUT/r02k_bench_${W}_g$N.err; echo "$W g$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02k_bench_${W}_g$N.json").read().strip().splitlines()[-1])
    print("$W g$N", "value=%.4g"%d["value"], "ms=%.4f"%d["ms_per_step"], "launches", d["gpu_launches"], "parity_ok=", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), "e2e=", (d.get("e2e") or {}).get("value"), d["clocks"].get("sm_mhz"), d["clocks"].get("samples"), d.get("slab_run"))
except Exception as e:
    print("$W g$N FAILED", e); print(open("$OUT/r02k_bench_${W}_g$N.err").read()[-2500:])
PY
}
run 8 lj1m 30
run 8 lj1m 100
run 4 lj1m 30
