"""The reference's own noise stream, generated on the device.

The reference draws every random number from a caller-owned ``numpy.random.Generator``
(``rgen.normal(...)`` at turtlemd/integrators.py:156-160 and :312; ``default_rng(seed)`` at :216).  When that
generator runs over ``PCG64`` -- what ``default_rng`` returns -- :class:`NumpyStream` continues it on the GPU:
``tmd_npstream_normals`` produces the very numbers ``rgen.standard_normal(count)`` would (PCG64 with jump-ahead, NumPy's
ziggurat, glibc's ``log1p`` for the tail; csrc/npy_normal.h), the integrator kernels consume them where they are, and the
host generator is moved to the position the device reached.  A stock script therefore keeps the reference's trajectory
without a host draw and a 48 MB upload per step.

Any other generator (``MT19937``, the reference tests' fake generators, ...) keeps the host-draw path of
:mod:`turtlemd_b200.integrators`.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _device, _lib

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())

_MASK = (1 << 64) - 1

ENABLED = True  # module switch: tests compare the device stream with the host-draw path
_warned = False


def usable(rgen) -> bool:
    """``rgen`` is a NumPy ``Generator`` over ``PCG64`` and this process can continue it on the device."""
    if not ENABLED or type(rgen) is not np.random.Generator:  # a subclass may override normal(): host path
        return False
    if type(rgen.bit_generator) is not np.random.PCG64:
        return False
    global _warned
    try:
        if _lib.load().tmd_npstream_libm_variant() < 0:
            if not _warned:
                _warned = True
                LOGGER.warning("this libm's log1p is neither glibc build restated in csrc/npy_normal.h: the NumPy generator "
                               "cannot be continued on the device bit for bit; Langevin noise is drawn on the host")
            return False
        return _lib.device_count() > 0
    except Exception:  # noqa: BLE001 - no library / no driver
        return False


def _words(state: int, inc: int) -> np.ndarray:
    return np.array([state >> 64, state & _MASK, inc >> 64, inc & _MASK], dtype=np.uint64)


def host_normals(rgen: np.random.Generator, count: int) -> np.ndarray:
    """``rgen.standard_normal(count)`` computed by the library's HOST twin of the device stream (same source), the
    generator advanced accordingly.  For tests and for C callers without NumPy; needs no device."""
    st = rgen.bit_generator.state
    words = _words(st["state"]["state"], st["state"]["inc"])
    out = np.empty(int(count))
    _lib.call("tmd_host_npstream_normals", _lib.hptr(words), _lib.hptr(out), int(count), None)
    st["state"]["state"] = (int(words[0]) << 64) | int(words[1])
    rgen.bit_generator.state = st
    return out


class NumpyStream:
    """Device continuation of one NumPy ``Generator`` (PCG64)."""

    def __init__(self, rgen: np.random.Generator):
        _device.require_cuda()
        self.rgen = rgen
        self._handle = C.c_void_p()
        _lib.call("tmd_npstream_create", C.byref(self._handle))
        self._known = None  # (state, inc) of the host generator when host and device last agreed
        self._buf = None
        self._open = False

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            try:
                _lib.call("tmd_npstream_destroy", handle)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass

    def begin(self) -> None:
        """Bring the device generator to the host generator's position (a no-op when nobody drew on the host since
        :meth:`end`)."""
        st = self.rgen.bit_generator.state["state"]
        now = (st["state"], st["inc"])
        if now != self._known:
            words = _words(*now)
            _lib.call("tmd_npstream_seed", self._handle, _lib.hptr(words), _device.stream_ptr())
            self._known = now
        self._open = True

    def normals(self, count: int):
        """The next ``count`` standard normals of the stream, as a flat device tensor (valid until the next call)."""
        if not self._open:
            raise RuntimeError("NumpyStream.normals outside begin() / end()")
        count = int(count)
        if self._buf is None or self._buf.numel() < count:
            self._buf = _device.empty(max(count, 1))
        _lib.call("tmd_npstream_normals", self._handle, _device.dptr(self._buf), count, _device.stream_ptr())
        return self._buf[:count]

    def end(self) -> None:
        """Wait for the stream and move the host generator to where the device stopped."""
        state = np.zeros(2, dtype=np.uint64)
        _lib.call("tmd_npstream_state", self._handle, _lib.hptr(state), None, None, _device.stream_ptr())
        st = self.rgen.bit_generator.state
        st["state"]["state"] = (int(state[0]) << 64) | int(state[1])
        self.rgen.bit_generator.state = st
        self._known = (st["state"]["state"], st["state"]["inc"])
        self._open = False

    def stats(self) -> tuple[int, int, int]:
        """(64-bit words consumed since the last seed, chunks walked again off the canonical chain by any pass, chunks
        whose normals were placed off the canonical chain)."""
        used, counters = C.c_uint64(0), np.zeros(2, dtype=np.uint64)
        _lib.call("tmd_npstream_state", self._handle, None, C.byref(used), _lib.hptr(counters), _device.stream_ptr())
        return used.value, int(counters[0]), int(counters[1])
