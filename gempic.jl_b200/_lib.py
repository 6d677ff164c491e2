"""ctypes binding of libgempic_b200.so (the C ABI of include/gempic_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or no B200 is visible the
import / first call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgempic_b200.so")

c_dp = C.POINTER(C.c_double)
c_handle = C.c_uint64


class GempicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[gempic_b200 status {code}] {msg}")
        self.code = code


class ArgumentError(GempicError, ValueError):
    """GEMPIC_EINVAL -- the reference throws ArgumentError here."""


class AssertionFailed(GempicError, AssertionError):
    """GEMPIC_EASSERT -- an @assert of the reference."""


FUNC1D = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
FUNC2D = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double, C.c_void_p)

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C gempic.jl_b200/csrc). gempic.jl_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.gempic_last_error.restype = C.c_char_p
    L.gempic_stream.restype = C.c_void_p
    L.gempic_launch_count.restype = C.c_int64
    L.gempic_launch_count.argtypes = [C.c_int]
    _lib = L
    return L


def check(rc: int):
    if rc == 0:
        return
    msg = load().gempic_last_error().decode(errors="replace")
    if rc == 1:
        raise ArgumentError(rc, msg)
    if rc == 2:
        raise AssertionFailed(rc, msg)
    raise GempicError(rc, msg)


def dptr(a):
    """double* of a contiguous float64 numpy array (None -> NULL)."""
    if a is None:
        return None
    import numpy as np

    assert isinstance(a, np.ndarray) and a.dtype == np.float64, "expected a float64 numpy array"
    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"], "expected a contiguous array"
    return a.ctypes.data_as(c_dp)


_initialised = False


def init(device: int | None = None):
    """Bind this process to a CUDA device (default: LOCAL_RANK or 0)."""
    global _initialised
    if _initialised:
        return
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    check(load().gempic_init(C.c_int(device)))
    _initialised = True
    _register_atexit()


_atexit_done = False


def _register_atexit():
    """finalise the library before the interpreter tears its objects down in arbitrary order (and, after
    init_devices, join the per-device worker threads)"""
    global _atexit_done
    if not _atexit_done:
        import atexit

        atexit.register(finalize)
        _atexit_done = True


def init_devices(devices):
    """ONE process driving several devices (gempic_init_devices): `devices` is a list of CUDA device ids or a count.
    Afterwards every object is replicated / sharded over them inside the library; the Python API is unchanged
    (ParticleGroup takes the global particle count)."""
    global _initialised
    if _initialised:
        raise GempicError(1, "the library is already initialised")
    ids = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
    arr = (C.c_int * len(ids))(*ids)
    check(load().gempic_init_devices(C.c_int(len(ids)), arr))
    _initialised = True
    _register_atexit()


def device_count() -> int:
    return int(load().gempic_device_count())


def finalize():
    global _initialised
    if _initialised:
        check(load().gempic_finalize())
        _initialised = False
