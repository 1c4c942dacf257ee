"""The potential plug-in interface (``turtlemd/potentials/potential.py:21-68``).

This ABC is the drop-in boundary of the force field: ``System`` only ever calls
``potential``, ``force`` and ``potential_and_force``.  The potentials of this
package additionally implement ``_launch(system, pos, force, vw, flags)``, which
enqueues their CUDA kernels on device-resident arrays; a plain host subclass
without ``_launch`` still works inside a ``System`` (its NumPy results are
uploaded and added).
"""
from __future__ import annotations

import logging
from abc import ABC, abstractmethod
from typing import Any

import numpy as np

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())


class Potential(ABC):
    """Base class of all potential functions."""

    desc: str
    dim: int
    params: Any

    def __init__(self, dim: int = 1, desc: str = ""):
        self.dim = dim
        self.desc = desc
        self.params = None

    def set_parameters(self, parameters: Any):
        """Default: log and ignore (potential.py:35-41)."""
        LOGGER.info(
            "Set parameters used, but it is not implemented by the"
            " potential - ignoring the given parameters: %s",
            parameters,
        )

    @abstractmethod
    def potential(self, system) -> float:
        """Potential energy."""

    @abstractmethod
    def force(self, system) -> tuple[np.ndarray, np.ndarray]:
        """Forces and virial."""

    def potential_and_force(self, system) -> tuple[float, np.ndarray, np.ndarray]:
        """Both; subclasses fuse them when that is cheaper (potential.py:51-61)."""
        vpot = self.potential(system)
        force, virial = self.force(system)
        return vpot, force, virial

    def __str__(self) -> str:
        lines = [f"Potential: {self.desc}"]
        lines.extend(f"{key}: {self.params[key]}" for key in sorted(self.params))
        return "\n".join(lines)
