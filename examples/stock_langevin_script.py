"""A Langevin run written the way a TurtleMD user writes it -- ``LangevinInertia(timestep, gamma, beta, seed=...)``, i.e. the
reference's own ``numpy.random.default_rng(seed)`` noise (turtlemd/integrators.py:216) -- with only the import changed.

The generator is continued ON THE DEVICE (PCG64 + NumPy's ziggurat, bit for bit: turtlemd_b200/npstream.py,
csrc/npy_normal.cu), so the run follows the trajectory the reference would compute under the same seed and still never
leaves HBM between steps.  ``--check`` repeats a short stretch with the draws made on the host like the reference makes them
and prints the difference (rounding level).

    python examples/stock_langevin_script.py [particles] [steps] [--check]
"""
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))  # run from a checkout

from turtlemd_b200 import npstream
from turtlemd_b200.integrators import LangevinInertia
from turtlemd_b200.potentials import DoubleWell
from turtlemd_b200.system import Box, Particles, System


def build(n):
    system = System(Box(periodic=[False]), Particles.from_arrays(np.full((n, 1), -1.0)), [DoubleWell(a=1.0, b=2.0, c=0.0)])
    system.potential_and_force()
    return system, LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, seed=7)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if len(args) > 0 else 20_000
    steps = int(args[1]) if len(args) > 1 else 500
    system, integrator = build(n)
    for block in range(5):
        integrator.run(system, steps // 5)
        x = system.particles.pos[:, 0]
        print(f"after {(block + 1) * (steps // 5):6d} steps: <x> = {x.mean():+.6f}  fraction in the right well = {(x > 0).mean():.4f}")
    print("the generator now stands where the reference's would:", integrator.rgen.bit_generator.state["state"]["state"])
    if "--check" in sys.argv:
        a, ia = build(n)
        ia.run(a, 20)
        npstream.ENABLED = False  # draw on the host, exactly like the reference
        b, ib = build(n)
        ib.run(b, 20)
        npstream.ENABLED = True
        print("device stream vs host draws after 20 steps: max |dx| =", float(np.max(np.abs(a.particles.pos - b.particles.pos))),
              " same generator state:", ia.rgen.bit_generator.state == ib.rgen.bit_generator.state)


if __name__ == "__main__":
    main()
