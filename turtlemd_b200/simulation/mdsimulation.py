"""The MD driver loop (API of ``turtlemd/simulation/mdsimulation.py``).

``step()`` / ``run()`` keep the reference's contract (step 0 only evaluates the
force field; ``steps=N`` yields N+1 systems; every stop condition is asked
before and inside each step).  ``run_batch`` is the device-resident fast path:
k integrator steps per host round trip, stop conditions polled between batches.
"""
from __future__ import annotations

import logging
from typing import TypedDict

from ..integrators import MDIntegrator
from ..system.system import System
from .stopping import MaxSteps, SoftExit, StopCondition

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())


class CycleNumber(TypedDict):
    step: int    # current step, counted from ``start``
    steps: int   # number of steps to do
    start: int   # first step number (for output)
    stepno: int  # steps done so far


class MDSimulation:
    """A system driven by an integrator until a stop condition fires."""

    def __init__(self, system: System, integrator: MDIntegrator, steps: int, start: int = 0,
                 stop_conditions: None | list[StopCondition] = None):
        self.system = system
        self.integrator = integrator
        self.steps = steps
        self.first_step = True
        self.cycle: CycleNumber = {"step": start, "steps": steps, "start": start, "stepno": 0}
        self.stop_conditions = [MaxSteps(), SoftExit()] if stop_conditions is None else stop_conditions
        if not any(isinstance(cond, SoftExit) for cond in self.stop_conditions):
            self.stop_conditions.append(SoftExit())  # there is always a way to stop a run from outside

    def stop(self) -> bool:
        return any(condition(self) for condition in self.stop_conditions)

    def step(self) -> System:
        """One step; the very first call only evaluates energies and forces (mdsimulation.py:69-81)."""
        if self.stop():
            return self.system
        if self.first_step:
            self.system.potential_and_force()
            self.first_step = False
            return self.system
        self.integrator.integration_step(self.system)
        self.cycle["step"] += 1
        self.cycle["stepno"] += 1
        return self.system

    def run(self):
        """Generator over the system after each step (mdsimulation.py:87-90)."""
        while not self.stop():
            yield self.step()

    def run_batch(self, batch: int = 100):
        """Like ``run`` but yields only every ``batch`` steps; in between nothing leaves the GPU.

        Uses the integrator's multi-step ``run(system, k)`` when it has one.  Stop conditions
        (``MaxSteps``, the ``EXIT`` file) are checked between batches.
        """
        multi = getattr(self.integrator, "run", None)
        while not self.stop():
            if self.first_step:
                self.system.evaluate(7)
                self.first_step = False
                yield self.system
                continue
            todo = min(batch, self.cycle["steps"] - self.cycle["stepno"])
            if todo <= 0:
                break
            if multi is not None:
                multi(self.system, todo)
            else:
                for _ in range(todo):
                    self.integrator.integration_step(self.system)
            self.cycle["step"] += todo
            self.cycle["stepno"] += todo
            yield self.system
