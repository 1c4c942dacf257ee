"""Simulation boxes (API of ``turtlemd/system/box.py``).

The pair kernels fuse the minimum-image rule of :meth:`Box.pbc_dist_matrix`
(box.py:384-403) and :meth:`Box.pbc_dist` (box.py:370-382); what they need from
a box is ``length``, the *stored* reciprocal ``ilength`` and ``periodic``
(:meth:`Box.to_abi`).  The methods below are the host-side helpers user
scripts call directly (e.g. ``pbc_wrap`` in the reference's movie example);
they are setup/analysis code, not the hot path.
"""
from __future__ import annotations

import itertools
import logging
from abc import ABC, abstractmethod

import numpy as np

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())


def guess_dimensionality(low=None, high=None, periodic=None) -> int:
    """Number of dimensions implied by whichever of low/high/periodic is given (box.py:12-34)."""
    sizes = [len(item) for item in (low, high, periodic) if item is not None]
    if not sizes:
        LOGGER.warning("Missing low/high/periodic parameters for box: assuming 1D")
        return 1
    if any(size != sizes[0] for size in sizes):
        LOGGER.error("Inconsistent box dimensions for low/high/periodic!")
        raise ValueError("Inconsistent box dimensions for low/high/periodic!")
    return sizes[0]


def cosine(angle: float) -> float:
    """cos of an angle in degrees, exactly 0 near 90 (box.py:37-51)."""
    return 0.0 if np.isclose(angle, 90) else np.cos(np.radians(angle))


def box_matrix_from_angles(length, alpha, beta, gamma, dim) -> np.ndarray:
    """Upper-triangular cell matrix, LAMMPS convention (box.py:54-89)."""
    cell = np.zeros((dim, dim))
    ca, cb, cg = cosine(alpha), cosine(beta), cosine(gamma)
    cell[0, 0] = length[0]
    cell[0, 1] = length[1] * cg
    cell[1, 1] = np.sqrt(length[1] ** 2 - cell[0, 1] ** 2)
    if dim > 2:
        cell[0, 2] = length[2] * cb
        cell[1, 2] = (length[1] * length[2] * ca - cell[0, 1] * cell[0, 2]) / cell[1, 1]
        cell[2, 2] = np.sqrt(length[2] ** 2 - cell[0, 2] ** 2 - cell[1, 2] ** 2)
    return cell


class BoxBase(ABC):
    """Common state of the boxes (box.py:92-174)."""

    def __init__(self, dim, periodic, low=None, high=None):
        self.dim = dim
        self.periodic = list(periodic) if periodic is not None else [True] * dim
        assert self.dim == len(self.periodic)
        # one degree of freedom is removed per periodic direction
        self.dof = np.array([1 if flag else 0 for flag in self.periodic])
        self.box_matrix = np.zeros((dim, dim))
        if low is None:
            self.low = np.array([0.0 if flag else -float("inf") for flag in self.periodic])
            LOGGER.warning("Set box low values: %s", self.low)
        else:
            self.low = np.asarray(low).astype(float)
        if high is None:
            self.high = np.array([1.0 if flag else float("inf") for flag in self.periodic])
            LOGGER.warning("Set box high values: %s", self.high)
        else:
            self.high = np.asarray(high).astype(float)
        self.length = self.high - self.low
        with np.errstate(divide="ignore"):
            self.ilength = 1.0 / self.length

    def volume(self) -> float:
        return np.linalg.det(self.box_matrix)

    @abstractmethod
    def pbc_wrap(self, pos: np.ndarray) -> np.ndarray:
        """Positions folded into the cell."""

    @abstractmethod
    def pbc_dist(self, distance: np.ndarray) -> np.ndarray:
        """Minimum image of one distance vector."""

    @abstractmethod
    def pbc_dist_matrix(self, distance: np.ndarray) -> np.ndarray:
        """Minimum image of many distance vectors."""


class Box(BoxBase):
    """Orthorhombic box (box.py:321-411): the one the CUDA kernels support."""

    def __init__(self, low=None, high=None, periodic=None):
        super().__init__(
            dim=guess_dimensionality(low=low, high=high, periodic=periodic),
            periodic=periodic, low=low, high=high,
        )
        self.box_matrix = np.diag(self.length.astype(float))

    def to_abi(self):
        """The ``tmd_box`` the kernels take."""
        from .. import _lib

        return _lib.make_box(self.dim, self.periodic, self.length, self.ilength)

    def pbc_wrap(self, pos: np.ndarray) -> np.ndarray:
        out = np.array(pos, dtype=float, copy=True)
        for k, flag in enumerate(self.periodic):
            if not flag:
                continue
            rel = pos[:, k] - self.low[k]
            outside = np.logical_or(rel < 0.0, rel >= self.length[k])
            folded = rel - np.floor(rel * self.ilength[k]) * self.length[k]
            out[:, k] = np.where(outside, folded, rel) + self.low[k]
        return out

    def pbc_dist(self, distance: np.ndarray) -> np.ndarray:
        out = np.array(distance, dtype=float, copy=True)
        for k, flag in enumerate(self.periodic):
            if flag and np.abs(distance[k]) > 0.5 * self.length[k]:  # strict '>' (box.py:376)
                out[k] = distance[k] - np.rint(distance[k] * self.ilength[k]) * self.length[k]
        return out

    def pbc_dist_matrix(self, distance: np.ndarray) -> np.ndarray:
        out = np.copy(distance)
        for k, flag in enumerate(self.periodic):
            if flag:
                column = out[:, k]
                far = np.where(np.abs(column) >= 0.5 * self.length[k])[0]  # '>=' (box.py:401)
                column[far] -= np.rint(column[far] * self.ilength[k]) * self.length[k]
        return out

    def __str__(self) -> str:
        return f"Hello, this is box and my matrix is:\n{self.box_matrix}\nPeriodic? {self.periodic}"


class TriclinicBox(BoxBase):
    """Triclinic cell (box.py:177-318).  ``LennardJonesCut`` and ``DoubleWellPair`` evaluate in it through
    ``tmd_lj_triclinic`` / ``tmd_dwpair_triclinic`` (all-pairs, exact image scan); the cell / neighbour-list
    kernels and the multi-GPU decompositions take ``Box`` only."""

    def __init__(self, low=None, high=None, periodic=None, alpha=None, beta=None, gamma=None):
        super().__init__(
            dim=guess_dimensionality(low=low, high=high, periodic=periodic),
            periodic=periodic, low=low, high=high,
        )
        self.alpha = 90 if alpha is None else alpha
        self.beta = 90 if beta is None else beta
        self.gamma = 90 if gamma is None else gamma
        self.box_matrix = np.zeros((self.dim, self.dim))
        self.box_matrix_inv = np.zeros((self.dim, self.dim))
        self.translations = np.zeros((3**self.dim, self.dim))
        skewed = not (self.alpha == self.beta == self.gamma == 90)
        if not (skewed and self.dim > 1):
            raise ValueError(
                "Not a valid triclinic box."
                "\n-At least one angle must be != 90."
                "\n-The dimensionality must be > 1."
            )
        self.update_box_matrix(self.length, self.alpha, self.beta, self.gamma)

    def to_abi(self):
        """The ``tmd_cell`` the triclinic pair kernel takes (``tmd_lj_triclinic``)."""
        from .. import _lib

        return _lib.make_cell(self.dim, self.box_matrix, self.box_matrix_inv, self.translations,
                              self.shortest_width_half)

    def update_box_matrix(self, length, alpha, beta, gamma):
        self.box_matrix = box_matrix_from_angles(length, alpha, beta, gamma, self.dim)
        self.box_matrix_inv = np.linalg.inv(self.box_matrix)
        self.translations = np.array(
            [self.box_matrix @ np.array(step) for step in itertools.product([-1, 0, 1], repeat=self.dim)]
        )
        self.shortest_width = np.min(np.diagonal(self.box_matrix))
        self.shortest_width_half = 0.5 * self.shortest_width

    def pbc_wrap(self, pos: np.ndarray) -> np.ndarray:
        frac = pos @ self.box_matrix_inv.T
        frac = frac - np.floor(frac)
        return frac @ self.box_matrix.T + self.low

    def pbc_dist(self, distance: np.ndarray) -> np.ndarray:
        """Nearest image after Mezei (1992): reduce, then scan the 3^dim neighbour cells."""
        dist = distance - self.box_matrix @ np.rint(self.box_matrix_inv @ distance)
        best = np.dot(dist, dist)
        if np.sqrt(best) < self.shortest_width_half:
            return dist
        images = dist + self.translations
        lengths = np.sum(images * images, axis=1)
        idx = np.argmin(lengths)
        return images[idx] if lengths[idx] < best else dist

    def pbc_dist_matrix(self, distance: np.ndarray) -> np.ndarray:
        return np.array([self.pbc_dist(row) for row in distance])

    def __str__(self) -> str:
        return (
            "Hello, this is triclinic box and my matrix "
            f"is:\n{self.box_matrix}Periodic? {self.periodic}"
        )
