"""Golden fixtures for the velocity set-up, generated from the live reference.

    python tests/golden/make_golden_maxwell.py

The reference's ``generate_maxwell_velocities`` / ``zero_momentum`` (turtlemd/system/particles.py:140-153,
:227-269) are fed the oracle's Philox twin as ``rgen``, so that the device path -- which generates the same
stream in a kernel -- can be compared number by number.  Same shim and helpers as make_golden.py.
"""
from __future__ import annotations

import numpy as np

import make_golden as mg  # noqa: F401  (puts the reference and the repo on sys.path)
from oracle.philox_ref import OraclePhiloxNormal
from turtlemd.system.particles import generate_maxwell_velocities, zero_momentum

CASES = [
    # tag, n, dim, temperature, boltzmann, dof, momentum, key, masses
    ("3d_500", 500, 3, 0.8, 1.0, [1.0, 1.0, 1.0], True, 11, "mixed"),
    ("1d_64", 64, 1, 2.0, 1.0, None, False, 5, "ones"),
    ("2d_301", 301, 2, 0.35, 0.5, [1.0, 0.0], True, 2024, "random"),
]


def main():
    for tag, n, dim, temperature, boltzmann, dof, momentum, key, masses in CASES:
        rng = np.random.default_rng(100 + n)
        if masses == "ones":
            mass = np.ones(n)
        elif masses == "mixed":
            mass = np.where(np.arange(n) % 3 == 0, 39.948, 1.0)
        else:
            mass = 0.5 + 4.0 * rng.random(n)
        part = mg.bulk_particles(rng.random((n, dim)), mass=mass)
        rgen = OraclePhiloxNormal(key=key)
        generate_maxwell_velocities(part, rgen, temperature=temperature, boltzmann=boltzmann, dof=dof,
                                    momentum=momentum)
        vel = part.vel.copy()
        # zero_momentum alone, with an axis mask, on velocities that carry a drift
        part.vel = vel + np.linspace(0.1, 0.3, dim)
        mask = [True, False, True][:dim]
        zero_momentum(part, dim=mask)
        mg.save(f"maxwell_{tag}.npz", mass=mass, key=key, temperature=temperature, boltzmann=boltzmann,
                dof=np.array([]) if dof is None else np.array(dof), momentum=momentum, vel=vel,
                drift=np.linspace(0.1, 0.3, dim), mask=np.array(mask), vel_masked=part.vel.copy(),
                draws=int(rgen.position))
        print(tag, "T check", 2.0 * 0.5 * np.sum(mass[:, None] * vel * vel) / (n * dim), "draws", rgen.position)


if __name__ == "__main__":
    main()
