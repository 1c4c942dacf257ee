set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/r02w_bench_default_g$N.json 2> $OUT/r02w_bench_default_g$N.err ) 2>&1 | grep real
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02w_bench_default_g$N.json").read().strip().splitlines()[-1])
    print("lj1m g$N", "value=%.4g"%d["value"], "ms=%.4f"%d["ms_per_step"], "parity_ok=", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), "e2e=", (d.get("e2e") or {}).get("value"), d["clocks"].get("sm_mhz"), d["clocks"].get("samples"), d.get("slab_run"))
    for c in d.get("configs", []): print("  ", c.get("workload"), c.get("value"), c.get("ms_per_step"), "parity_ok=", (c.get("parity") or {}).get("ok"), c.get("error"), (c.get("parallelism") or "")[:40])
except Exception as e:
    print("FAILED", e); print(open("$OUT/r02w_bench_default_g$N.err").read()[-2500:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 profiles/profile_slabrun.py 30 graph > $OUT/r02w_prof_g${N}_graph.log 2>&1; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM\|Warn\|warn" $OUT/r02w_prof_g${N}_graph.log | tail -14
