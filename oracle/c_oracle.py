"""ctypes face of oracle/c/em_oracle.c (TEST / CPU-BASELINE INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) and build:
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_SO)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_threads.argtypes = [C.c_int]
        _lib.oracle_set_threads.restype = None
        _lib.oracle_flat_fit.restype = None
        _lib.oracle_flat_fit.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
    return _lib


def set_threads(n):
    load().oracle_set_threads(int(n))


def num_threads():
    return int(load().oracle_num_threads())


def flat_fit(X, mu0, iters, sigma0_sq=1.0):
    """-> (weights [J], mu [J,3], cov [J,3,3], ll [iters]) in float64, all host threads."""
    lib = load()
    X = np.ascontiguousarray(X, np.float32)
    mu0 = np.ascontiguousarray(mu0, np.float32)
    J = mu0.shape[0]
    w = np.zeros(J)
    mu = np.zeros((J, 3))
    cov = np.zeros((J, 3, 3))
    ll = np.zeros(iters)
    lib.oracle_flat_fit(X.ctypes.data, X.shape[0], J, mu0.ctypes.data, float(sigma0_sq), int(iters), w.ctypes.data, mu.ctypes.data,
                        cov.ctypes.data, ll.ctypes.data)
    return w, mu, cov, ll
