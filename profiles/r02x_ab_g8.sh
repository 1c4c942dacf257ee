set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
for F in 1 0; do for D in 1 0; do
  TMD_SR_FUSED_BARRIER=$F TMD_SR_DYNAMIC=$D TMD_BENCH_CONFIGS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 60 --warmup 5 --quick > $OUT/r02x_ab_f${F}_d${D}_g$N.json 2> $OUT/r02x_ab.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02x_ab_f${F}_d${D}_g$N.json").read().strip().splitlines()[-1])
    print("fused=$F dynamic=$D g$N", "value=%.4g"%d["value"], "ms=%.4f"%d["ms_per_step"])
except Exception as e:
    print("FAILED", e); print(open("$OUT/r02x_ab.err").read()[-1500:])
PY
done; done
