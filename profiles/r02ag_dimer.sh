set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_gates.py -m gpu -x -q 2>&1 | tail -12
TMD_BENCH_CONFIGS=0 timeout 300 python bench.py --workload dimer256k --quick > $OUT/r02ag_bench_dimer.json 2> $OUT/r02ag.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02ag_bench_dimer.json").read().strip().splitlines()[-1])
print("dimer256k", d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parallelism"])
PY
