"""turtlemd_b200: the TurtleMD hot path (pair potentials + integrators) on B200 GPUs.

Same module layout and class names as ``turtlemd`` (infretis/turtlemd 2023.3.1),
so a script ports by changing the top-level package name.  All arithmetic of
potentials and integrators runs in hand-written sm_100a CUDA kernels behind the
C ABI of ``include/turtlemd_b200.h``; there is no CPU fallback.
"""
from .version import __version__

__all__ = ["__version__"]
