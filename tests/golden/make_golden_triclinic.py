"""Golden fixtures for LennardJonesCut in a TriclinicBox, generated from the live reference.

    python tests/golden/make_golden_triclinic.py

potentials/lennardjones.py:145-273 with system/box.py:183-318 (Mezei reduction + 3^dim image scan).  The
cells are strongly skewed and the cut-off exceeds half the shortest cell width in one case, so that both
branches of TriclinicBox.pbc_dist decide pairs inside the cut-off.  Same shim and helpers as make_golden.py.
"""
from __future__ import annotations

import numpy as np

import make_golden as mg  # noqa: F401  (puts the reference and the repo on sys.path)
from turtlemd.potentials import DoubleWellPair, LennardJonesCut
from turtlemd.system import System
from turtlemd.system.box import TriclinicBox

CASES = [
    # tag, n, high, angles (alpha, beta, gamma), rcut scale, seed
    ("3d_150", 150, [10.5, 9.75, 9.0], (55.0, 60.0, 50.0), 1.0, 7),
    ("3d_90_longcut", 90, [9.0, 9.0, 8.25], (70.0, 65.0, 75.0), 1.7, 8),
    ("2d_80", 80, [13.5, 12.0], (None, None, 55.0), 1.0, 9),
]
PARAMS = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.1, "epsilon": 0.8, "rcut": 2.2}}


def main():
    for tag, n, high, (alpha, beta, gamma), rscale, seed in CASES:
        dim = len(high)
        rng = np.random.default_rng(seed)
        box = TriclinicBox(high=high, alpha=alpha, beta=beta, gamma=gamma)
        # well separated points: a jittered grid in fractional coordinates, then some atoms pushed into
        # neighbouring images (positions are never wrapped by the reference)
        per = int(np.ceil(n ** (1.0 / dim)))
        grid = np.stack(np.meshgrid(*[np.arange(per)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:n]
        frac = (grid + 0.5 + 0.25 * (2.0 * rng.random((n, dim)) - 1.0)) / per
        frac += rng.integers(-1, 2, size=(n, dim)) * (rng.random((n, 1)) < 0.2)
        pos = frac @ box.box_matrix.T
        ptype = (rng.random(n) < 0.3).astype(int)
        params = {k: dict(v, rcut=v["rcut"] * rscale) for k, v in PARAMS.items()}
        pot = LennardJonesCut(dim=dim, shift=True, mixing="geometric")
        pot.set_parameters(params)
        system = System(box, mg.bulk_particles(pos, ptype=ptype), [pot])
        v, f, w = pot.potential_and_force(system)
        v_only = pot.potential(system)
        f_only, w_only = pot.force(system)
        assert np.array_equal(f, f_only)
        mg.save(f"triclinic_lj_{tag}.npz", pos=pos, ptype=ptype, high=np.array(high, dtype=float),
                angles=np.array([np.nan if a is None else a for a in (alpha, beta, gamma)]),
                rscale=rscale, v=v, f=f, w=w, v_only=v_only, w_only=w_only,
                box_matrix=box.box_matrix, half_width=box.shortest_width_half)
        # DoubleWellPair in the same skewed cell (potentials/well.py:230-286 with TriclinicBox.pbc_dist)
        if n <= 100:
            for types in ((1, 1), (0, 1)):
                dw = DoubleWellPair(types=types, dim=dim)
                dw.set_parameters({"rzero": 1.2, "width": 0.3, "height": 5.0})
                sys_dw = System(box, mg.bulk_particles(pos, ptype=ptype), [dw])
                fd, wd = dw.force(sys_dw)
                mg.save(f"triclinic_dwpair_{tag}_types{types[0]}{types[1]}.npz", pos=pos, ptype=ptype,
                        high=np.array(high, dtype=float),
                        angles=np.array([np.nan if a is None else a for a in (alpha, beta, gamma)]),
                        types=np.array(types), rzero=1.2, width=0.3, height=5.0, v=dw.potential(sys_dw), f=fd, w=wd)
        near = np.sum(np.abs(f) > 0) // dim
        print(tag, "V", v, "atoms with force", near, "half width", box.shortest_width_half,
              "max rc", max(p["rcut"] for p in params.values()))


if __name__ == "__main__":
    main()
