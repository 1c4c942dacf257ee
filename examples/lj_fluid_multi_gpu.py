"""The Lennard-Jones fluid of examples/lj_fluid.py on several B200s of one node: one process per GPU, the system cut
into one slab of cell layers per GPU (parallel.SlabRun; LangevinInertia with --langevin).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/lj_fluid_multi_gpu.py [cells_per_side] [steps] [--langevin]

Inside a run nothing leaves the GPUs and no collective library is called: the traversal kernel of a rank stores the
records of its two boundary layers straight into its neighbours' memory, one barrier kernel per step agrees on list
rebuilds.  NCCL only sets the job up and gathers the state at the end.
"""
import os
import pathlib
import sys
import time

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))  # run from a checkout

import torch
import torch.distributed as dist

from turtlemd_b200.integrators import LangevinInertia, VelocityVerlet
from turtlemd_b200.parallel import SlabRun
from turtlemd_b200.philox import PhiloxNormal
from turtlemd_b200.potentials import LennardJonesCut
from turtlemd_b200.system import Box, Particles, System
from turtlemd_b200.system.particles import generate_maxwell_velocities
from turtlemd_b200.tools import generate_lattice


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep = int(args[0]) if len(args) > 0 else 63
    steps = int(args[1]) if len(args) > 1 else 200
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()

    # every rank builds the same system (same seeds): the slab run distributes it
    pos, size = generate_lattice("fcc", [rep, rep, rep], density=0.8)
    particles = Particles.from_arrays(pos)
    generate_maxwell_velocities(particles, rgen=PhiloxNormal(key=1), temperature=0.8)
    potential = LennardJonesCut(dim=3, shift=True, method="nlist")
    potential.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
    system = System(Box(low=size[:, 0], high=size[:, 1]), particles, [potential])
    system.potential_and_force()  # a run starts from valid forces
    if "--langevin" in sys.argv:
        integrator = LangevinInertia(timestep=0.005, gamma=0.3, beta=1.0 / 0.8, rgen=PhiloxNormal(key=2))
    else:
        integrator = VelocityVerlet(0.005)

    run = SlabRun(system, integrator)
    run.prepare()
    run.capture()  # pairs of steps replayed from the library's step graph
    run.run(10)    # warm-up: module load, first graph
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    run.run(steps)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    v_pot, virial = run.reduce_observables()
    x, v, f = run.gather_state()
    if rank == 0:
        kin = 0.5 * float(np.sum(v * v))
        status = run.status()
        print(f"{len(pos)} atoms on {world} GPUs ({status['layers_per_rank']} of {status['layers']} cell layers each), {steps} steps "
              f"in {sec * 1e3:.1f} ms = {len(pos) * steps / sec:.3e} atom-steps/s; {status['builds']} list rebuilds")
        print(f"E/N = {(kin + v_pot) / len(pos):+.6f}  T = {2.0 * kin / (3 * len(pos) - 3):.4f}  "
              f"P = {(np.trace(virial) + 2.0 * kin) / (3.0 * system.box.volume()):.4f}")
    run.release()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
