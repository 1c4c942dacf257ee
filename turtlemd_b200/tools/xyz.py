"""Reading xyz files (API of ``turtlemd/tools/xyz.py``).  Host I/O, outside the hot path."""
from __future__ import annotations

import logging
import pathlib
from collections.abc import Iterator
from dataclasses import dataclass, field

import numpy as np

from ..system.particles import Particles

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())


@dataclass
class Snapshot:
    """One frame of an xyz file."""

    natoms: int = 0
    comment: str = ""
    atoms: list[str] = field(default_factory=list)
    _xyz: list[list[float]] = field(default_factory=list)
    xyz: np.ndarray = field(default_factory=lambda: np.zeros(3))


def read_xyz_file(filename: str | pathlib.Path) -> Iterator[Snapshot]:
    """Yield the frames of a (multi-frame) xyz file (xyz.py:26-62)."""
    frame, remaining = None, 0
    with open(filename) as handle:
        for line in handle:
            if remaining == 0:
                if frame is not None:
                    frame.xyz = np.array(frame._xyz)
                    yield frame
                    frame = None
                try:
                    remaining = int(line.strip())
                except ValueError:
                    LOGGER.error("Could not read the number of atoms (first line) in file %s!", filename)
                    raise
                header = True
                continue
            if header:
                frame = Snapshot(natoms=remaining, comment=line.strip())
                header = False
                continue
            remaining -= 1
            columns = line.strip().split()
            frame.atoms.append(columns[0])
            frame._xyz.append([float(c) for c in columns[1:]])
    if frame is not None:
        frame.xyz = np.array(frame._xyz)
        yield frame


def particles_from_xyz_file(filename: str | pathlib.Path, dim: int = 3,
                            masses: dict[str, float] | None = None) -> Particles:
    """Particles from the first frame; extra columns are velocities, then forces (xyz.py:65-100)."""
    masses = masses or {}
    frame = next(read_xyz_file(filename))
    particles = Particles(dim=dim)
    ptypes = {atom: idx for idx, atom in enumerate(set(frame.atoms))}
    for atom, row in zip(frame.atoms, frame.xyz):
        vel = force = None
        if len(row) > dim:
            vel = row[dim : 2 * dim]
            vel = vel if len(vel) == dim else None
            force = row[2 * dim : 3 * dim]
            force = force if len(force) == dim else None
        particles.add_particle(pos=row[:dim], vel=vel, force=force, mass=masses.get(atom, 1.0),
                               name=atom, ptype=ptypes.get(atom, -1))
    return particles
