"""Host twin of the in-kernel noise stream ("tmd-normal-v1").

:class:`PhiloxNormal` is a duck-typed random generator (``normal(loc, scale,
size)``, which is all the reference integrators ask of ``rgen``,
``turtlemd/integrators.py:156-160`` and ``:312``).  Its numbers come from the
*same C source* the CUDA kernels compile (``csrc/philox_normal.h``, host build
inside ``libturtlemd_b200.so``): normal number ``q`` of the stream with key
``k`` is a pure function of ``(k, q)``, bit-identical on host and device.

* handed to the *reference* integrators it makes them consume exactly the draws
  the kernels generate, so trajectories can be compared step by step;
* handed to this package's Langevin integrators it selects the in-kernel stream
  (no host draw, no upload) and is advanced as if it had been drawn from.

Raw words are those of ``numpy.random.Philox(key=k)``.
"""
from __future__ import annotations

import numpy as np

from . import _lib


class PhiloxNormal:
    """Counter-based normal stream; ``position`` = index of the next normal."""

    def __init__(self, key: int = 0, position: int = 0):
        self.key = int(key) & 0xFFFFFFFFFFFFFFFF
        self.position = int(position)

    def normal(self, loc=0.0, scale=1.0, size=None):
        """``loc + scale * z`` for the next ``prod(size)`` normals, C order."""
        if size is None:
            shape = ()
        elif np.iterable(size):
            shape = tuple(int(s) for s in size)
        else:
            shape = (int(size),)
        count = int(np.prod(shape)) if shape else 1
        z = self.peek(count)
        self.position += count
        if not shape:
            return loc + scale * z[0]
        return loc + scale * z.reshape(shape)

    def peek(self, count: int, offset: int = 0) -> np.ndarray:
        """Normals ``position+offset .. +count`` without advancing."""
        out = np.empty(count, dtype=np.float64)
        _lib.call("tmd_host_philox_normal_fill", _lib.hptr(out), count, self.key, self.position + offset)
        return out

    def raw(self, count: int, first: int = 0) -> np.ndarray:
        """Raw uint64 words ``first .. first+count-1`` (== ``numpy.random.Philox(key).random_raw``)."""
        out = np.empty(count, dtype=np.uint64)
        _lib.call("tmd_host_philox_raw_fill", _lib.hptr(out), count, self.key, int(first))
        return out

    def advance(self, count: int) -> None:
        self.position += int(count)

    def __repr__(self) -> str:
        return f"PhiloxNormal(key={self.key}, position={self.position})"
