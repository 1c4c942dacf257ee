"""Potentials: ``LennardJonesCut``, ``DoubleWell``, ``DoubleWellPair`` (as ``turtlemd.potentials``)."""
from .lennardjones import LennardJonesCut
from .well import DoubleWell, DoubleWellPair, RectangularWell

__all__ = ["LennardJonesCut", "DoubleWell", "DoubleWellPair", "RectangularWell"]
