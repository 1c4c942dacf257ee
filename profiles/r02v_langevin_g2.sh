set -u
OUT=gpurun_out; mkdir -p $OUT
for G in eager graph; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_check.py slabrun 24 41 $G langevin > $OUT/r02v_check_$G.log 2>&1; echo "check $G rc=$?"; grep MULTI_GPU_CHECK $OUT/r02v_check_$G.log || tail -25 $OUT/r02v_check_$G.log
done
N=2
for W in lj1m_langevin; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --workload $W > $OUT/r02v_bench_${W}_g$N.json 2> $OUT/r02v_bench_${W}_g$N.err; echo "$W rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02v_bench_${W}_g$N.json").read().strip().splitlines()[-1])
    print("$W", "value=%.4g"%d["value"], "ms=%.4f"%d["ms_per_step"], "launches", d["gpu_launches"], "\n parity=", d.get("parity"), "\n e2e=", (d.get("e2e") or {}).get("value"), "\n", d["config"]["parallelism"][:80], "\n", d.get("slab_run"))
except Exception as e:
    print("$W FAILED", e); print(open("$OUT/r02v_bench_${W}_g$N.err").read()[-2500:])
PY
done
