set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
for G in eager graph; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_check.py slabrun 24 41 $G dimer > $OUT/r02ah_check_$G.log 2>&1; echo "check $G rc=$?"; grep MULTI_GPU_CHECK $OUT/r02ah_check_$G.log || tail -25 $OUT/r02ah_check_$G.log
done
for W in dimer256k; do
  TMD_BENCH_CONFIGS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --workload $W > $OUT/r02ah_bench_${W}_g$N.json 2> $OUT/r02ah_bench_${W}_g$N.err; echo "$W rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02ah_bench_${W}_g$N.json").read().strip().splitlines()[-1])
    print("$W", "value=%.4g"%d["value"], "ms=%.4f"%d["ms_per_step"], "launches", d["gpu_launches"], "\n parity=", d.get("parity"), "\n", d["config"]["parallelism"][:60], "\n", d.get("slab_run"))
except Exception as e:
    print("$W FAILED", e); print(open("$OUT/r02ah_bench_${W}_g$N.err").read()[-2500:])
PY
done
