"""ctypes binding of ``libturtlemd_b200.so`` (the C ABI in ``include/turtlemd_b200.h``).

The library is the product: if it is missing this module tries ONE in-tree
build (``make`` in ``csrc/``) and otherwise raises -- there is no Python or
NumPy fallback for any kernel.  Status codes become exceptions here.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import shutil
import subprocess

import numpy as np

CSRC = pathlib.Path(__file__).resolve().parent / "csrc"
LIBPATH = CSRC / "libturtlemd_b200.so"

WANT_V, WANT_F, WANT_W, ACCUMULATE, V_STANDALONE = 1, 2, 4, 8, 16
MAX_TYPES = 8

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_NOMEM = range(6)


class TurtleCudaError(RuntimeError):
    """A C-ABI call failed; ``status`` is the tmd_status code."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class NoDeviceError(TurtleCudaError):
    """No CUDA device: the package has no CPU path for the hot path."""


class Box(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("periodic", C.c_int32 * 3),
        ("length", C.c_double * 3),
        ("ilength", C.c_double * 3),
    ]


class Cell(C.Structure):
    """``tmd_cell``: what ``TriclinicBox.pbc_dist`` needs (system/box.py:183-318)."""

    _fields_ = [
        ("dim", C.c_int32), ("reserved", C.c_int32),
        ("matrix", C.c_double * 9), ("inverse", C.c_double * 9),
        ("translations", C.c_double * 81), ("half_width", C.c_double),
    ]


class Rng(C.Structure):
    """``tmd_rng``; ``step_counter`` / ``step_stride``: optional device-side step counter (CUDA graph replays)."""

    _fields_ = [("key", C.c_uint64), ("offset", C.c_uint64), ("step_counter", C.c_void_p),
                ("step_stride", C.c_uint64)]


class LangevinConsts(C.Structure):
    _fields_ = [(name, C.c_double) for name in (
        "c0", "a1", "a2f", "b1f", "b2f", "dt", "beta", "beta_gamma", "kfac",
        "one_minus_e2", "one_minus_e_sq")]


class DwPairTerm(C.Structure):
    """``tmd_dwpair_term``: a DoubleWellPair as the second potential of a resident run."""

    _fields_ = [("idx0", C.c_void_p), ("n0", C.c_int32), ("idx1", C.c_void_p), ("n1", C.c_int32),
                ("same_type", C.c_int32), ("rwidth", C.c_double), ("width2", C.c_double), ("height", C.c_double)]


class NlistBuffers(C.Structure):
    """``tmd_nlist_buffers``: device pointers of a neighbour-list handle (slab path)."""

    _fields_ = [
        ("records", C.c_void_p), ("build_records", C.c_void_p), ("order", C.c_void_p),
        ("slot_of", C.c_void_p), ("image", C.c_void_p), ("flags", C.c_void_p),
        ("cells", C.c_int32 * 3), ("half_width", C.c_int32), ("capacity", C.c_int32),
    ]


_P = C.c_void_p
_I64, _I32, _U32, _U64, _D = C.c_int64, C.c_int32, C.c_uint32, C.c_uint64, C.c_double
_BOX, _RNG, _LC = C.POINTER(Box), C.POINTER(Rng), C.POINTER(LangevinConsts)
_NLB = C.POINTER(NlistBuffers)
_CELL = C.POINTER(Cell)

# name -> argtypes; every function returns int (tmd_status) unless listed in _NON_STATUS
PROTOTYPES = {
    "tmd_abi_version": [],
    "tmd_last_error": [],
    "tmd_status_string": [C.c_int],
    "tmd_device_count": [C.POINTER(C.c_int)],
    "tmd_set_device": [C.c_int],
    "tmd_device_info": [C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int)],
    "tmd_stream_synchronize": [_P],
    "tmd_launch_count": [C.POINTER(_I64)],
    "tmd_malloc": [C.POINTER(_P), C.c_size_t],
    "tmd_free": [_P],
    "tmd_memset_async": [_P, C.c_int, C.c_size_t, _P],
    "tmd_memcpy_h2d_async": [_P, _P, C.c_size_t, _P],
    "tmd_memcpy_d2h_async": [_P, _P, C.c_size_t, _P],
    "tmd_memcpy_d2d_async": [_P, _P, C.c_size_t, _P],
    "tmd_host_register": [_P, C.c_size_t],
    "tmd_host_unregister": [_P],
    "tmd_ipc_get_handle": [_P, _P],
    "tmd_ipc_open_handle": [_P, C.POINTER(_P)],
    "tmd_ipc_close_handle": [_P],
    "tmd_lj_allpairs": [_P, _P, _I64, _BOX, _P, _I32, _U32, _I64, _I64, _P, _P, _P],
    "tmd_lj_cells_usable": [_I64, _BOX, _P, _I32, C.POINTER(C.c_int)],
    "tmd_lj_cells": [_P, _P, _I64, _BOX, _P, _I32, _U32, _I64, _I64, _P, _P, _P],
    "tmd_nlist_create": [C.POINTER(_P), _D],
    "tmd_nlist_destroy": [_P],
    "tmd_lj_nlist_usable": [_I64, _BOX, _P, _I32, _D, C.POINTER(C.c_int)],
    "tmd_nlist_stats": [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32)],
    "tmd_lj_nlist": [_P, _P, _P, _I64, _BOX, _P, _I32, _U32, _I64, _I64, _P, _P, _P],
    "tmd_nlist_run_vv": [_P, _P, _P, _P, _P, _P, _I64, _BOX, _P, _I32, _D, _I32, _P, _P],
    "tmd_nlist_run_vv_dwpair": [_P, _P, _P, _P, _P, _P, _I64, _BOX, _P, _I32, _D, _I32, C.POINTER(DwPairTerm), _P, _P],
    "tmd_nlist_set_timing": [_P, _I32],
    "tmd_nlist_timing": [_P, C.POINTER(_D), C.POINTER(_I64)],
    "tmd_nlist_bind_records": [_P, _P],
    "tmd_nlist_build_slots": [_P, _P, _P, _I64, _BOX, _P, _I32, _I64, _I64, _I32, _P],
    "tmd_nlist_force_slots": [_P, _U32, _P, _P, _P],
    "tmd_nlist_view": [_P, _NLB],
    "tmd_slab_vv_kick_drift": [_P, _P, _P, _P, _P, _I64, _I64, _I32, _D, _D, _P, _P],
    "tmd_slab_langevin_inertia_pre": [_P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _LC, _RNG, _D, _P, _P],
    "tmd_slab_unsort": [_P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P],
    "tmd_slab_sort": [_P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P],
    "tmd_slabrun_create": [C.POINTER(_P), _I32, _I32, _I64, _BOX, _P, _I32, _D, _D],
    "tmd_slabrun_destroy": [_P],
    "tmd_slabrun_window": [_P, C.POINTER(_P), C.POINTER(C.c_size_t)],
    "tmd_slabrun_connect": [_P, C.POINTER(_P)],
    "tmd_slabrun_load": [_P, _P, _P, _P, _P, _P, _P],
    "tmd_slabrun_set_langevin": [_P, _LC, _U64, _U64],
    "tmd_nlist_run_langevin": [_P, _P, _P, _P, _P, _P, _I64, _BOX, _P, _I32, _LC, _RNG, _I32, _P, _P],
    "tmd_slabrun_set_dwpair": [_P, _P, _I32, _P, _I32, _I32, _D, _D, _D, _BOX],
    "tmd_slabrun_start": [_P, _P],
    "tmd_slabrun_step": [_P, _I32, _P],
    "tmd_slabrun_run": [_P, _I32, _I32, C.POINTER(_I64), _P],
    "tmd_slabrun_phase": [_P, _I32, _I32, _P],
    "tmd_slabrun_export": [_P, _P, _P, _P, _P, _P],
    "tmd_slabrun_status": [_P, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32)],
    "tmd_doublewell": [_P, _I64, _I32, _D, _D, _D, _U32, _P, _P, _P],
    "tmd_dwpair": [_P, _I64, _BOX, _P, _I32, _P, _I32, _I32, _D, _D, _D, _U32, _P, _P, _P],
    "tmd_vv_kick_drift": [_P, _P, _P, _P, _I64, _I32, _D, _P],
    "tmd_vv_kick": [_P, _P, _P, _I64, _I32, _D, _P],
    "tmd_vv_kick_kick_drift": [_P, _P, _P, _P, _I64, _I32, _D, _P],
    "tmd_verlet_step": [_P, _P, _P, _P, _P, _I64, _I32, _D, _I32, _P],
    "tmd_langevin_overdamped": [_P, _P, _P, _P, _I64, _I32, _D, _D, _D, _RNG, _I64, _P, _P],
    "tmd_langevin_inertia_constants": [_D, _D, _D, _LC],
    "tmd_langevin_inertia_pre": [_P, _P, _P, _P, _I64, _I32, _LC, _RNG, _I64, _P, _P, _P],
    "tmd_langevin_inertia_post_pre": [_P, _P, _P, _P, _I64, _I32, _LC, _RNG, _I64, _P],
    "tmd_langevin_inertia_post": [_P, _P, _P, _I64, _I32, _LC, _P],
    "tmd_langevin_overdamped_normals": [_P, _P, _P, _P, _I64, _I32, _D, _D, _D, _P, _P],
    "tmd_langevin_inertia_pre_normals": [_P, _P, _P, _P, _I64, _I32, _LC, _P, _P, _I32, _P],
    "tmd_npstream_libm_variant": [],
    "tmd_npstream_create": [C.POINTER(_P)],
    "tmd_npstream_destroy": [_P],
    "tmd_npstream_seed": [_P, _P, _P],
    "tmd_npstream_normals": [_P, _P, _I64, _P],
    "tmd_npstream_state": [_P, _P, C.POINTER(_U64), _P, _P],
    "tmd_host_npstream_normals": [_P, _P, _I64, C.POINTER(_U64)],
    "tmd_langevin_inertia_doublewell": [_P, _P, _P, _P, _I64, _I32, _D, _D, _D, _LC, _RNG, _I64, _I64, _I32, _P, _P],
    "tmd_counter_add": [_P, _U64, _P],
    "tmd_philox_normal_fill": [_P, _I64, _RNG, _P],
    "tmd_host_philox_normal_fill": [_P, _I64, _U64, _U64],
    "tmd_host_philox_raw_fill": [_P, _I64, _U64, _U64],
    "tmd_kinetic_tensor": [_P, _P, _I64, _I32, _P, _P],
    "tmd_lj_triclinic": [_P, _P, _I64, _CELL, _P, _I32, _U32, _I64, _I64, _P, _P, _P],
    "tmd_host_lj_triclinic": [_P, _P, _I64, _CELL, _P, _I32, _U32, _P, _P],
    "tmd_dwpair_triclinic": [_P, _I64, _CELL, _P, _I32, _P, _I32, _I32, _D, _D, _D, _U32, _P, _P, _P],
    "tmd_host_dwpair_triclinic": [_P, _P, _I64, _CELL, _I64, _I64, _D, _D, _D, _U32, _P, _P],
    "tmd_maxwell_draw": [_P, _P, _I64, _I32, _RNG, _I64, _P],
    "tmd_momentum": [_P, _P, _I64, _I32, _P, _P],
    "tmd_vel_shift_scale": [_P, _I64, _I32, _P, _I32, _D, _P],
    "tmd_host_lj": [_P, _P, _I64, _BOX, _P, _I32, _U32, _I32, _P, _P],
    "tmd_host_doublewell": [_P, _I64, _I32, _D, _D, _D, _U32, _P, _P],
    "tmd_host_dwpair": [_P, _P, _I64, _BOX, _I64, _I64, _D, _D, _D, _U32, _P, _P],
    "tmd_host_vv_lj_step": [_P, _P, _P, _P, _P, _I64, _BOX, _P, _I32, _I32, _D, _P],
    "tmd_bench_dfma": [_I32, C.POINTER(_D), C.POINTER(_D)],
    "tmd_bench_copy": [C.c_size_t, _I32, C.POINTER(_D), C.POINTER(_D)],
}
_NON_STATUS = {"tmd_npstream_libm_variant": C.c_int, "tmd_abi_version": C.c_int, "tmd_last_error": C.c_char_p, "tmd_status_string": C.c_char_p}

_lib = None


def build(force: bool = False) -> pathlib.Path:
    """Compile the CUDA library in-tree for sm_100a (``make`` in ``csrc/``)."""
    if shutil.which("nvcc") is None:
        raise RuntimeError("nvcc not found: cannot build libturtlemd_b200.so")
    if force:
        subprocess.run(["make", "-C", str(CSRC), "clean"], check=True, capture_output=True)
    proc = subprocess.run(["make", "-C", str(CSRC), "-j8"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libturtlemd_b200.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    return LIBPATH


def load() -> C.CDLL:
    """Load (building once if needed) the library and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIBPATH.exists():
        build()
    lib = C.CDLL(str(LIBPATH))
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.argtypes = argtypes
        fn.restype = _NON_STATUS.get(name, C.c_int)
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == OK:
        return
    lib = load()
    text = lib.tmd_last_error().decode() or lib.tmd_status_string(status).decode()
    if status == ERR_NO_DEVICE:
        raise NoDeviceError(status, text)
    if status == ERR_INVALID:
        raise ValueError(text)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    raise TurtleCudaError(status, text)


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))


_graph_launches = 0  # kernel launches replayed from CUDA graphs (the library only sees the recording)


def launch_count() -> int:
    """Kernels launched by the library so far in this process (graph replays included)."""
    n = C.c_int64(0)
    call("tmd_launch_count", C.byref(n))
    return n.value + _graph_launches


def add_launches(count: int) -> None:
    global _graph_launches
    _graph_launches += int(count)


def device_count() -> int:
    n = C.c_int(0)
    call("tmd_device_count", C.byref(n))
    return n.value


def make_cell(dim: int, matrix, inverse, translations, half_width: float) -> Cell:
    """``tmd_cell`` from the attributes ``TriclinicBox`` stores (system/box.py:244-258)."""
    cell = Cell()
    cell.dim = int(dim)
    for r in range(dim):
        for k in range(dim):
            cell.matrix[r * 3 + k] = float(matrix[r][k])
            cell.inverse[r * 3 + k] = float(inverse[r][k])
    for g in range(3 ** dim):
        for k in range(dim):
            cell.translations[g * 3 + k] = float(translations[g][k])
    cell.half_width = float(half_width)
    return cell


def make_box(dim: int, periodic, length, ilength) -> Box:
    """``tmd_box`` from the attributes ``Box`` stores (system/box.py:140-158)."""
    box = Box()
    box.dim = int(dim)
    for k in range(3):
        on = k < dim and bool(periodic[k])
        box.periodic[k] = 1 if on else 0
        box.length[k] = float(length[k]) if on else 0.0
        box.ilength[k] = float(ilength[k]) if on else 0.0
    return box


def hptr(array: np.ndarray) -> C.c_void_p:
    """Pointer to a C-contiguous NumPy array."""
    return C.c_void_p(array.ctypes.data)
