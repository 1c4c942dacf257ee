"""Stop conditions of a simulation (API of ``turtlemd/simulation/stopping.py``)."""
from __future__ import annotations

import logging
import pathlib
from abc import ABC, abstractmethod

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())


class StopCondition(ABC):
    """Callable predicate on a simulation."""

    @abstractmethod
    def stop(self, simulation) -> bool:
        """True when the simulation should end."""

    def __call__(self, simulation) -> bool:
        return self.stop(simulation)


class SoftExit(StopCondition):
    """Stop when a file named ``EXIT`` exists in ``simulation.exe_dir`` (default: cwd) (stopping.py:31-41)."""

    def stop(self, simulation) -> bool:
        folder = getattr(simulation, "exe_dir", pathlib.Path())
        if (folder / "EXIT").is_file():
            LOGGER.info("Found exit file - will exit at the next step.")
            return True
        return False


class MaxSteps(StopCondition):
    """Stop after ``cycle["steps"]`` steps (stopping.py:44-49)."""

    def stop(self, simulation) -> bool:
        return simulation.cycle["stepno"] >= simulation.cycle["steps"]
