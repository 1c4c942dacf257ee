"""System objects: ``Box``, ``Particles``, ``System`` (as ``turtlemd.system``)."""
from .box import Box, TriclinicBox
from .particles import Particles
from .system import System

__all__ = ["Box", "TriclinicBox", "Particles", "System"]
