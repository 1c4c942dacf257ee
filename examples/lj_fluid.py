"""A Lennard-Jones fluid with TurtleMD's API on a B200 (the workload of the reference's examples/lennard.py,
written for this package; same classes, same calls).

    python examples/lj_fluid.py [cells_per_side] [steps]

cells_per_side = 4 gives the reference's 256-atom argon box; 63 gives the N = 1 000 188 benchmark system.
"""
import pathlib
import sys
import time

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))  # run from a checkout

from turtlemd_b200.integrators import VelocityVerlet
from turtlemd_b200.philox import PhiloxNormal
from turtlemd_b200.potentials import LennardJonesCut
from turtlemd_b200.simulation import MDSimulation
from turtlemd_b200.system import Box, Particles, System
from turtlemd_b200.system.particles import generate_maxwell_velocities
from turtlemd_b200.tools import generate_lattice


def main():
    rep = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    pos, size = generate_lattice("fcc", [rep, rep, rep], density=0.8)
    box = Box(low=size[:, 0], high=size[:, 1])
    particles = Particles.from_arrays(pos)  # bulk constructor; add_particle() works too (O(N^2) for N atoms)
    generate_maxwell_velocities(particles, rgen=PhiloxNormal(key=1), temperature=0.8)  # drawn on the device
    potential = LennardJonesCut(dim=3, shift=True)  # all-pairs below 4096 atoms, Verlet neighbour list above
    potential.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
    system = System(box, particles, [potential])
    simulation = MDSimulation(system, VelocityVerlet(0.005), steps=steps)
    t0 = time.perf_counter()
    for frame in simulation.run_batch(batch=max(1, steps // 10)):  # nothing leaves the GPU inside a batch
        th = frame.thermo()
        print(f"step {simulation.cycle['stepno']:6d}  T = {th['temperature']:.4f}  E/N = {th['ekin'] + th['vpot']:+.6f}  "
              f"P = {th['pressure']:.4f}")
    sec = time.perf_counter() - t0
    print(f"{len(pos)} atoms, {steps} steps in {sec:.2f} s = {len(pos) * steps / sec:.3e} atom-steps/s (wall clock, "
          "with the thermo read-backs)")
    assert np.isfinite(system.particles.v_pot)


if __name__ == "__main__":
    main()
