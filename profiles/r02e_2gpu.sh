set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_real_ranks" > $OUT/r02e_2rank_tests.log 2>&1; echo "2-rank tests rc=$?"; tail -8 $OUT/r02e_2rank_tests.log
for W in lj1m lj1m_langevin dimer256k replicas; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload $W > $OUT/r02e_bench_${W}_g2.json 2> $OUT/r02e_bench_${W}_g2.err; echo "$W rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02e_bench_${W}_g2.json").read().strip().splitlines()[-1])
    print("$W", "value=%.4g"%d["value"], "ms=%.4f"%d["ms_per_step"], "parity=", d.get("parity"), "e2e=", (d.get("e2e") or {}).get("value"), d["clocks"].get("sm_mhz"), d["clocks"].get("samples"))
except Exception as e:
    print("$W FAILED", e); print(open("$OUT/r02e_bench_${W}_g2.err").read()[-1500:])
PY
done
