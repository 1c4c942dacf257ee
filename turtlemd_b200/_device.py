"""Device residency for the particle arrays.

PyTorch is used for exactly three things here: allocating HBM
(``torch.empty(..., device="cuda")``), naming the current CUDA stream and
host<->device copies.  Every kernel is launched through the C ABI with raw
pointers (``tensor.data_ptr()``).

:class:`Mirror` pairs one host ``ndarray`` (what the reference API hands out)
with one device tensor and tracks which side is current:

* reading ``particles.pos`` downloads if the device copy is newer and then --
  because the caller may write into the array it was handed
  (``particles.pos[0] = x`` is legal reference usage) -- marks the device copy
  stale, so the next kernel re-uploads.  Conservative, never wrong.
* kernels ask for :meth:`Mirror.device` (upload if stale) and call
  :meth:`Mirror.written_on_device` after writing.

A loop that only calls ``integrator(system)`` therefore never leaves HBM.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_torch = None


def torch():
    """Import torch lazily (host-only use of the package must not need CUDA)."""
    global _torch
    if _torch is None:
        import torch as _t

        _torch = _t
    return _torch


def require_cuda() -> None:
    """Raise :class:`NoDeviceError` unless a CUDA device and the CUDA library are usable."""
    _lib.load()
    if _lib.device_count() < 1 or not torch().cuda.is_available():
        raise _lib.NoDeviceError(
            _lib.ERR_NO_DEVICE,
            "turtlemd_b200 needs a CUDA device (sm_100a); there is no CPU fallback for "
            "potentials or integrators",
        )


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch().cuda.current_stream().cuda_stream)


def dptr(tensor) -> C.c_void_p:
    return C.c_void_p(tensor.data_ptr())


def synchronize() -> None:
    torch().cuda.current_stream().synchronize()


def empty(shape, dtype="float64"):
    t = torch()
    return t.empty(shape, dtype=getattr(t, dtype), device="cuda")


def zeros(shape, dtype="float64"):
    t = torch()
    return t.zeros(shape, dtype=getattr(t, dtype), device="cuda")


def pinned_array(shape) -> np.ndarray:
    """float64 host array in page-locked memory (fast H2D/D2H); falls back to pageable without CUDA."""
    t = torch()
    if not t.cuda.is_available():
        return np.zeros(shape)
    return t.zeros(int(np.prod(shape)), dtype=t.float64, pin_memory=True).numpy().reshape(shape)


_side = None


def _side_stream():
    """One extra stream per process for downloads that overlap kernels (Mirror.prefetch_host)."""
    global _side
    if _side is None:
        _side = torch().cuda.Stream()
    return _side


def to_device(array: np.ndarray, dtype=None):
    """Host array -> new device tensor (flattened, contiguous)."""
    t = torch()
    host = np.ascontiguousarray(array if dtype is None else array.astype(dtype))
    return t.from_numpy(host.reshape(-1)).to("cuda", non_blocking=False)


class Mirror:
    """One logical float64 array with a host (NumPy) and a device (torch) copy."""

    __slots__ = ("_host", "_dev", "host_valid", "dev_valid", "_host_read", "_wants_host", "_prefetch")

    def __init__(self, host: np.ndarray):
        self._host = host
        self._dev = None
        self.host_valid = True
        self.dev_valid = False
        self._host_read = False   # the user has looked at the host array since the last device write
        self._wants_host = False  # ... and had done so for the value before that one: worth prefetching
        self._prefetch = None     # event of an asynchronous download in flight

    # -- host side ---------------------------------------------------------------------------
    def host(self) -> np.ndarray:
        """The array for the USER: current, and assumed modified after this call."""
        self.refresh_host()
        self.dev_valid = False
        self._host_read = True
        return self._host

    def peek(self) -> np.ndarray:
        """Current host array for read-only internal use (device copy stays valid)."""
        self.refresh_host()
        return self._host

    def prefetch_host(self) -> None:
        """Start downloading the device copy on a side stream, behind everything queued so far on the
        current stream.  Integrators call this right after the kernel that wrote the array when other
        kernels that only READ it follow (positions: the whole force evaluation), so that a user who reads
        the array after every step does not wait for the copy afterwards.  Only when that user exists (the
        previous value was read), and only into a page-locked array of the right shape."""
        if self.host_valid or not self._wants_host or self._dev is None or self._prefetch is not None:
            return
        t = torch()
        host = self._host
        if not (host.size == self._dev.numel() and host.dtype == np.float64 and host.flags.writeable
                and host.flags.c_contiguous):
            return
        target = t.from_numpy(host.reshape(-1))
        if not target.is_pinned():
            return
        side = _side_stream()
        side.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(side):
            target.copy_(self._dev, non_blocking=True)
            self._prefetch = side.record_event()

    def _finish_prefetch(self, on_device: bool) -> None:
        """Wait for a download in flight: the host (default) or, on_device, the current stream."""
        if self._prefetch is not None:
            if on_device:
                torch().cuda.current_stream().wait_event(self._prefetch)
            else:
                self._prefetch.synchronize()

    def refresh_host(self) -> None:
        if self.host_valid:
            return
        if self._prefetch is not None:  # the download was started when the array was written
            self._prefetch.synchronize()
            self._prefetch = None
            self.host_valid = True
            return
        t = torch()
        host = self._host
        if (host.size == self._dev.numel() and host.dtype == np.float64 and host.flags.writeable
                and host.flags.c_contiguous):
            # straight into the array users may hold a reference to (DMA when it is page-locked)
            t.from_numpy(host.reshape(-1)).copy_(self._dev)
        else:
            self._host = self._dev.cpu().numpy().reshape(host.shape if host.size == self._dev.numel() else (-1,))
        self.host_valid = True

    def set_host(self, array) -> None:
        """Rebind, as ``particles.vel = array`` does in the reference."""
        self._finish_prefetch(on_device=False)  # it may be writing into the very array that comes back
        self._prefetch = None
        self._host = np.asarray(array)
        self.host_valid = True
        self.dev_valid = False

    # -- device side -------------------------------------------------------------------------
    def device(self, read_only: bool = False):
        """Current device tensor (flat float64); uploads when the host copy is newer.  ``read_only``: the
        caller's kernels only read it (they may run beside a download in flight); anybody else may be
        about to overwrite the tensor, so the current stream first waits for that download."""
        if not read_only:
            self._finish_prefetch(on_device=True)
        if not self.dev_valid:
            host = np.ascontiguousarray(self._host, dtype=np.float64)
            if self._dev is None or self._dev.numel() != host.size:
                self._dev = to_device(host)
            else:
                self._dev.copy_(torch().from_numpy(host.reshape(-1)), non_blocking=False)
            self.dev_valid = True
        return self._dev

    def device_uninitialised(self, size: int):
        """Device tensor that is about to be overwritten completely (no upload)."""
        if self._dev is None or self._dev.numel() != size:
            self._dev = empty(size)
        return self._dev

    def written_on_device(self, shape=None) -> None:
        self.dev_valid = True
        self.host_valid = False
        self._prefetch = None  # (already waited for in device(): what it fetched is stale now)
        self._wants_host, self._host_read = self._host_read, False
        if shape is not None and tuple(self._host.shape) != tuple(shape):
            self._host = np.zeros(shape)

    @property
    def shape(self):
        return self._host.shape
