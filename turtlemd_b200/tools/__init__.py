"""Setup helpers (as ``turtlemd.tools``)."""
from .tools import generate_lattice
from .xyz import particles_from_xyz_file, read_xyz_file

__all__ = ["generate_lattice", "particles_from_xyz_file", "read_xyz_file"]
