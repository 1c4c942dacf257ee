"""An ensemble of independent double-well Langevin replicas (the workload of the reference's
examples/doublewell.ipynb, every replica one 1-D particle), written for this package.

    python examples/doublewell_replicas.py [replicas] [steps]
"""
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))  # run from a checkout

from turtlemd_b200.integrators import LangevinInertia
from turtlemd_b200.philox import PhiloxNormal
from turtlemd_b200.potentials import DoubleWell
from turtlemd_b200.system import Box, Particles, System


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    system = System(Box(periodic=[False]), Particles.from_arrays(np.full((n, 1), -1.0)), [DoubleWell(a=1.0, b=2.0, c=0.0)])
    system.potential_and_force()
    integrator = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, rgen=PhiloxNormal(key=0))
    for block in range(10):
        integrator.run(system, steps // 10)  # all steps of a block in one kernel; replica i draws from stream position i
        x = system.particles.pos[:, 0]
        print(f"after {(block + 1) * (steps // 10):6d} steps: <x> = {x.mean():+.4f}  fraction in the right well = {(x > 0).mean():.4f}")


if __name__ == "__main__":
    main()
