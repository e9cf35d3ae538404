"""TEST INFRASTRUCTURE ONLY (oracle) -- ctypes access to oracle/_build/liboracle.so."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_poisson_batch.restype = C.c_long
        _lib.oracle_poisson_batch.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_counts_ext_batch.restype = C.c_long
        _lib.oracle_counts_ext_batch.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                                 C.c_void_p]
        _lib.oracle_poisson_stream.restype = C.c_long
        _lib.oracle_poisson_stream.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    return _lib


def poisson_batch(lam, uniforms):
    """lam: (...,) float64; uniforms: (..., draws).  Returns (counts int32, used int32)."""
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    assert u.shape[:-1] == lam.shape
    counts = np.empty(lam.shape, dtype=np.int32)
    used = np.empty(lam.shape, dtype=np.int32)
    lib().oracle_poisson_batch(lam.size, lam.ctypes.data, u.ctypes.data, u.shape[-1], counts.ctypes.data,
                               used.ctypes.data)
    return counts, used


def poisson_stream(lam, uniforms):
    lam = np.ascontiguousarray(lam, dtype=np.float64).ravel()
    u = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
    counts = np.empty(lam.size, dtype=np.int64)
    used = lib().oracle_poisson_stream(lam.size, lam.ctypes.data, u.ctypes.data, u.size, counts.ctypes.data)
    return counts, used


def counts_ext_batch(lam, uniforms, nb_phi=0.0, dropout_p=0.0):
    """The read-out's extension draws (negative binomial via Gamma-Poisson, Bernoulli dropout) over explicit uniforms;
    with both switches at 0 this is poisson_batch.  Returns (counts int32, used int32)."""
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    assert u.shape[:-1] == lam.shape
    counts = np.empty(lam.shape, dtype=np.int32)
    used = np.empty(lam.shape, dtype=np.int32)
    lib().oracle_counts_ext_batch(lam.size, lam.ctypes.data, u.ctypes.data, u.shape[-1], float(nb_phi), float(dropout_p),
                                  counts.ctypes.data, used.ctypes.data)
    return counts, used
