"""numpy facade standing in for CuPy so the reference's Python files import unmodified.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). The reference calls
cupy.get_array_module / cupy.dot / cupy.clip(a, a_min=...) / cupy.asnumpy /
cupy.cuda.Stream.null.synchronize (gmm_waymo/src/gmm_impl.py:29,45,86;
gmmreg_gpu/gmm_impl.py:51) and cp.random / cp.power / cp.copy
(hgmm/hgmm_cupy_cpu_working.py:97,124-125,158).
"""
import numpy as _np
from numpy import *  # noqa: F401,F403

float32 = _np.float32
random = _np.random


def get_array_module(*_a):
    return _np


def asnumpy(x):
    return _np.asarray(x)


def asarray(x, *a, **k):
    return _np.asarray(x, *a, **k)


def clip(a, a_min=None, a_max=None):
    return _np.clip(a, a_min, a_max)


class _Null:
    @staticmethod
    def synchronize():
        return None


class _Stream:
    null = _Null()


class cuda:  # noqa: N801
    Stream = _Stream
