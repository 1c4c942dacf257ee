"""One rank of an 8-way SLAB decomposition of the N = 1M fluid, run alone on one GPU (no exchange:
halo records simply stay where they were at the last build, which changes the physics slightly but
not the work), with the pieces of a step timed separately by CUDA events.

    python profiles/emulate_slab_rank.py [world] [rank] [steps]

Under ncu --metrics gpu__time_duration.sum the same run gives the per-kernel times."""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch

import bench
from turtlemd_b200.integrators import VelocityVerlet
from turtlemd_b200.parallel import ShardPlan, SlabDecomposition
from turtlemd_b200.potentials import LennardJonesCut
from turtlemd_b200.system import Box, Particles, System

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
pos, vel, size = bench.lj_inputs(63)
pot = LennardJonesCut(dim=3, shift=True, method="nlist")
pot.set_parameters(bench.SINGLE)
system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos, vel=vel), [pot])
dec = SlabDecomposition(system, VelocityVerlet(0.005), plan=ShardPlan(len(pos), world, rank))
dec.prepare()


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for _ in range(3):
    dec.step_move()
    dec.step_force()
t_move = timed(dec.step_move, steps)
t_force = timed(dec.step_force, steps)
t_rebuild = timed(dec.rebuild, 5)
dec.capture(warm_steps=2)
rebuilds0 = dec.rebuilds
t_step = timed(dec.step, steps)
print(f"world={world} rank={rank} halo={dec.halo} own={dec.plan.count}: move {t_move:.1f} us, force+kick {t_force:.1f} us, "
      f"rebuild (unsort, bin, lists, sort, reach check) {t_rebuild:.1f} us; graphed step with host flag read "
      f"{t_step:.1f} us ({dec.rebuilds - rebuilds0} rebuilds in {steps} steps)")
