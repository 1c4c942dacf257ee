"""Simulation driver (as ``turtlemd.simulation``)."""
from .mdsimulation import MDSimulation

__all__ = ["MDSimulation"]
