"""One rank of an 8-way atom decomposition of the N = 1M fluid, run alone on one GPU (no collectives:
positions of the other ranks simply stay where they are), so that its kernels can be timed with
ncu --metrics gpu__time_duration.sum.   python profiles/emulate_rank.py [world] [rank] [steps]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from turtlemd_b200.integrators import VelocityVerlet
from turtlemd_b200.parallel import AtomDecomposition, ShardPlan
from turtlemd_b200.potentials import LennardJonesCut
from turtlemd_b200.system import Box, Particles, System

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
pos, vel, size = bench.lj_inputs(63)
pot = LennardJonesCut(dim=3, shift=True, method="nlist")
pot.set_parameters(bench.SINGLE)
system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos, vel=vel), [pot])
dec = AtomDecomposition(system, VelocityVerlet(0.005), plan=ShardPlan(len(pos), world, rank))
dec.prepare()
for _ in range(3):
    dec.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    dec.step()
e1.record()
torch.cuda.synchronize()
print(f"world={world} rank={rank}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us per step without the exchange", pot.nlist_stats())
