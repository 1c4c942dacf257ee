"""Round-2 golden fixtures from the live reference (same recipe as make_golden.py: run HERE, in the build
container, where /root/reference is mounted; the GPU box only reads the .npz files).

    python tests/golden/make_golden_r2.py [name ...]

  traj_vv_lj_864_hot.npz      864-atom LJ fluid at T ~ 6, VelocityVerlet(0.004), frames 0 / 4 / 8 / 12 of a 12-step
                              MDSimulation run.  Large enough (L = 10.26 > 3 x 2.8) for the neighbour-list path
                              and hot enough that the list is rebuilt inside the run: pins the fused resident
                              run (tmd_nlist_run_vv) directly against the reference.
  traj_vv_dimer_864.npz       config #5 in miniature on the neighbour-list path: LJ (rc = 2.5, two types with
                              identical parameters) + DoubleWellPair on two bonded atoms, 10 steps.
  traj_langevin_inertia_lj_108_default_rng.npz  a stock LangevinInertia(seed=3) script over a 108-atom LJ fluid, 10 steps.
  langevin_dist_default_rng.npz  x, v of 20000 independent 1-D DoubleWell particles after 500 LangevinInertia steps
                              under the reference's OWN default generator (default_rng(seed), PCG64 + ziggurat):
                              the sample the "matching in distribution" gate compares the in-kernel Philox
                              stream against (integrators.py:216, :284-297).
"""
from __future__ import annotations

import sys

import numpy as np

from make_golden import (Box, DoubleWell, DoubleWellPair, LangevinInertia, LennardJonesCut, MDSimulation, System,
                         VelocityVerlet, bulk_particles, lj_fluid, save)


def traj_frames(name, system, integrator, steps, keep, extra):
    frames = {k: [] for k in ("pos", "vel", "force", "virial", "v_pot")}
    p = system.particles
    init = dict(pos0=p.pos.copy(), vel0=p.vel.copy(), mass=p.mass.copy(), ptype=p.ptype.copy())
    sim = MDSimulation(system, integrator, steps)
    for frame, sysi in enumerate(sim.run()):
        if frame in keep:
            q = sysi.particles
            frames["pos"].append(q.pos.copy())
            frames["vel"].append(q.vel.copy())
            frames["force"].append(q.force.copy())
            frames["virial"].append(q.virial.copy())
            frames["v_pot"].append(float(q.v_pot))
    out = {k: np.array(v) for k, v in frames.items()}
    out.update(init)
    out.update(extra)
    out["frames"] = np.array(sorted(keep))
    save(name, **out)


def vv_lj_864_hot():
    pos, size = lj_fluid([6] * 3, seed=61)
    rng = np.random.default_rng(62)
    vel = rng.standard_normal(pos.shape) * 2.4
    vel -= vel.mean(axis=0)
    part = bulk_particles(pos, vel=vel)
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[pot])
    traj_frames("traj_vv_lj_864_hot.npz", system, VelocityVerlet(0.004), 12, {0, 4, 8, 12},
                dict(low=box.low, high=box.high, dt=0.004))


def vv_dimer_864():
    pos, size = lj_fluid([6] * 3, seed=67)
    ptype = np.zeros(len(pos), dtype=int)
    first = len(pos) // 2
    near = int(np.argsort(np.sum((pos - pos[first]) ** 2, axis=1))[1])
    ptype[[first, near]] = 1
    rng = np.random.default_rng(68)
    vel = rng.standard_normal(pos.shape) * 1.5
    vel -= vel.mean(axis=0)
    part = bulk_particles(pos, vel=vel, ptype=ptype)
    one = {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}
    lj = LennardJonesCut(dim=3, shift=True)
    lj.set_parameters({0: dict(one), 1: dict(one)})
    dw = DoubleWellPair(types=(1, 1))
    dw.set_parameters({"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25})
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[lj, dw])
    traj_frames("traj_vv_dimer_864.npz", system, VelocityVerlet(0.004), 10, {0, 5, 10},
                dict(low=box.low, high=box.high, dt=0.004))


def langevin_dist():
    """The reference under its own default RNG.  Its per-particle Python loop makes 20000 particles x 500 steps
    take a few minutes; the seed is the reference's constructor argument."""
    n, steps = 20000, 500
    part = bulk_particles(np.full((n, 1), -1.0))
    system = System(box=Box(periodic=[False]), particles=part, potentials=[DoubleWell(a=1.0, b=2.0, c=0.0)])
    integ = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, seed=7)
    system.potential_and_force()
    for _ in range(steps):
        integ(system)
    save("langevin_dist_default_rng.npz", pos=system.particles.pos.copy(), vel=system.particles.vel.copy(),
         n=n, steps=steps, dt=0.002, gamma=0.3, beta=1.0 / 0.07, seed=7, a=1.0, b=2.0, c=0.0, x0=-1.0)


def langevin_lj_default_rng():
    """config #4 in miniature exactly as a stock script writes it: LangevinInertia(timestep, gamma, beta, seed=3) -- the
    reference's own default generator -- over a 108-atom LJ fluid, 10 bare integrator calls."""
    from make_golden import generate_maxwell_velocities, run_traj

    pos, size = lj_fluid([3] * 3, seed=71)
    part = bulk_particles(pos)
    generate_maxwell_velocities(part, rgen=np.random.default_rng(9), temperature=0.8)
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[pot])
    system.potential_and_force()
    integ = LangevinInertia(timestep=0.005, gamma=0.3, beta=1.0 / 0.8, seed=3)
    run_traj("traj_langevin_inertia_lj_108_default_rng.npz", system, integ, 10, first_eval=False,
             extra=dict(low=box.low, high=box.high, dt=0.005, gamma=0.3, beta=1.0 / 0.8, seed=3))


CASES = {"traj_vv_lj_864_hot.npz": vv_lj_864_hot, "traj_vv_dimer_864.npz": vv_dimer_864,
         "langevin_dist_default_rng.npz": langevin_dist,
         "traj_langevin_inertia_lj_108_default_rng.npz": langevin_lj_default_rng}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
