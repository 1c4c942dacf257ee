set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py tests/test_gpu_gates.py -m gpu -x -q -k "langevin or Langevin" 2>&1 | tail -15
TMD_BENCH_CONFIGS=0 timeout 300 python bench.py --workload lj1m_langevin --quick > $OUT/r02aa_bench_langevin.json 2> $OUT/r02aa.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02aa_bench_langevin.json").read().strip().splitlines()[-1])
print("lj1m_langevin", d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["parallelism"])
PY
