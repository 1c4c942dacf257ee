"""All ranks of an N-way resident slab run on ONE GPU, phase by phase (parallel.SlabRunEmulation), for per-kernel
timings of one rank's work under `ncu --metrics gpu__time_duration.sum` (ncu cannot wrap a multi-rank job):

    python profiles/emulate_slabrun.py [world] [steps] [rep]
"""
import sys
import pathlib

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

import torch

import bench
from turtlemd_b200.parallel import SlabRunEmulation

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 63
cfg = dict(rep=rep, method="nlist", dt=0.005)
pos, vel, size, n, describe, make = bench.build_lj(cfg)
system, integ, pot = make()
system.evaluate(7)
emu = SlabRunEmulation([system] * world, integ)
emu.run(5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
emu.run(steps)
e1.record()
torch.cuda.synchronize()
print(f"world={world} n={n} steps={steps}: {e0.elapsed_time(e1) / steps / world * 1e3:.1f} us per rank-step (serialised on one GPU)",
      emu.status()[0])
emu.release()
