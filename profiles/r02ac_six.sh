set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_slabrun.py -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import torch, bench
import turtlemd_b200.integrators as I
I.RESIDENT_LANGEVIN = True
cfg = bench.WORKLOADS["lj1m_langevin"]
pos, vel, size, n, describe, make = bench.build_lj(cfg)
system, integ, pot = make()
system.evaluate(7)
integ.run(system, 5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); integ.run(system, 60); e1.record(); torch.cuda.synchronize()
print("resident langevin ms/step", e0.elapsed_time(e1) / 60)
PY
