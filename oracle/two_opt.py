"""ORACLE (test infrastructure): ctypes front end of oracle/two_opt_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True)
        h = C.CDLL(_SO)
        h.oracle_batched_two_opt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64]
        h.oracle_nls.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int]
        h.oracle_numpy_pairwise_sum.argtypes = [C.c_void_p, C.c_int]
        h.oracle_numpy_pairwise_sum.restype = C.c_float
        h.oracle_tour_cost_numpy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        h.oracle_tour_cost_numpy.restype = C.c_float
        _lib = h
    return _lib


def batched_two_opt(dist, tours, max_iterations):
    """tsp_nls/two_opt.py:41-49.  dist [n,n] float32, tours [count,n] -> new uint16 array."""
    d = np.ascontiguousarray(dist, dtype=np.float32)
    t = np.ascontiguousarray(tours).astype(np.uint16).copy()
    lib().oracle_batched_two_opt(d.ctypes.data, d.shape[0], t.ctypes.data, t.shape[0], int(max_iterations))
    return t


def nls(dist, heu_dist, tours, maxt, T_nls=10, T_p=20):
    """tsp_nls/aco.py:241-258 on [count,n] tours."""
    d = np.ascontiguousarray(dist, dtype=np.float32)
    h = np.ascontiguousarray(heu_dist, dtype=np.float32)
    t = np.ascontiguousarray(tours).astype(np.uint16).copy()
    lib().oracle_nls(d.ctypes.data, h.ctypes.data, d.shape[0], t.ctypes.data, t.shape[0], int(maxt), int(T_nls), int(T_p))
    return t


def numpy_pairwise_sum(row):
    r = np.ascontiguousarray(row, dtype=np.float32)
    return np.float32(lib().oracle_numpy_pairwise_sum(r.ctypes.data, r.shape[0]))
