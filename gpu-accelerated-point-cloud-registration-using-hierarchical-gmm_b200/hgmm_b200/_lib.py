"""ctypes binding of libhgmm.so (C ABI: include/hgmm.h).

The product path has NO CPU fallback: if the shared library is missing or no CUDA device is
present, importing the library / creating an Engine raises.  Nothing here imports `oracle`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HGMM_LIB_PATH") or os.path.join(_HERE, "libhgmm.so")      # override: A/B builds of the library

HGMM_OK = 0
MEM_HOST, MEM_DEVICE = 0, 1
COV_FULL, COV_DIAG, COV_SPHERICAL = 0, 1, 2
FLAVOR_CPP, FLAVOR_PY, FLAVOR_PY_OLD = 0, 1, 2
LL_LEVEL, LL_ESTEP = 0, 1
SOLVER_TWIST_LSTSQ, SOLVER_PROCRUSTES = 0, 1

COV_TYPES = {"full": COV_FULL, "diag": COV_DIAG, "spherical": COV_SPHERICAL}
SOLVERS = {"twist_lstsq": SOLVER_TWIST_LSTSQ, "procrustes_svd": SOLVER_PROCRUSTES, "procrustes": SOLVER_PROCRUSTES}


class FlatConfig(C.Structure):
    _fields_ = [("n_components", C.c_int32), ("cov_type", C.c_int32), ("flavor", C.c_int32), ("max_iter", C.c_int32),
                ("tol", C.c_float), ("sigma_bug", C.c_int32), ("tile_points", C.c_int32), ("reserved", C.c_int32)]


class TreeConfig(C.Structure):
    _fields_ = [("max_level", C.c_int32), ("ll_mode", C.c_int32), ("ls", C.c_float), ("ld", C.c_float), ("sig2", C.c_float),
                ("max_iters_per_level", C.c_int32), ("chunk_points", C.c_int32), ("reserved", C.c_int32),
                ("prune_lambda_c", C.c_float), ("prune_min_points", C.c_float)]


class RegConfig(C.Structure):
    _fields_ = [("solver", C.c_int32), ("maxiter", C.c_int32), ("tol", C.c_float), ("lambda_c", C.c_float)]


class HgmmError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/hgmm.h declares
_VP = C.c_void_p
SIGNATURES = {
    "hgmm_create": (C.c_int, [C.POINTER(_VP), C.c_int, _VP]),
    "hgmm_destroy": (C.c_int, [_VP]),
    "hgmm_last_error": (C.c_char_p, [_VP]),
    "hgmm_version": (C.c_char_p, []),
    "hgmm_launch_count": (C.c_int64, [_VP]),
    "hgmm_set_points": (C.c_int, [_VP, _VP, C.c_int64, C.c_int]),
    "hgmm_total_points": (C.c_int64, [_VP]),
    "hgmm_declare_total_points": (C.c_int, [_VP, C.c_int64]),
    "hgmm_fit_flat": (C.c_int, [_VP, C.POINTER(FlatConfig), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "hgmm_predict_flat": (C.c_int, [_VP, _VP, C.c_int64, C.c_int, _VP]),
    "hgmm_tree_total_nodes": (C.c_int64, [C.c_int32]),
    "hgmm_fit_tree": (C.c_int, [_VP, C.POINTER(TreeConfig), _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "hgmm_tree_set_model": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP]),
    "hgmm_reg_set_target": (C.c_int, [_VP, _VP, C.c_int64, C.c_int]),
    "hgmm_reg_estep": (C.c_int, [_VP, _VP, _VP, C.c_float, _VP, _VP, _VP]),
    "hgmm_reg_mstep": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP]),
    "hgmm_register_tree": (C.c_int, [_VP, C.POINTER(RegConfig), _VP, _VP, _VP, _VP, _VP]),
    "hgmm_register_flat": (C.c_int, [_VP, C.POINTER(RegConfig), _VP, _VP, _VP, _VP, _VP]),
    "hgmm_l2_set_mixtures": (C.c_int, [_VP, _VP, _VP, C.c_int32, _VP, _VP, C.c_int32]),
    "hgmm_l2_cost_grad": (C.c_int, [_VP, _VP, C.c_double, _VP, _VP]),
    "hgmm_l2_optimize": (C.c_int, [_VP, _VP, C.c_double, C.c_int32, C.c_double, _VP, _VP, _VP, _VP]),
    "hgmm_fill_vbo": (C.c_int, [_VP, _VP, _VP, C.c_float, _VP, _VP]),
    "hgmm_comm_unique_id": (C.c_int, [_VP]),
    "hgmm_comm_init": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "hgmm_comm_destroy": (C.c_int, [_VP]),
    "hgmm_io_read_ply": (C.c_int, [C.c_char_p, C.c_int32, _VP, C.c_int64, _VP]),
    "hgmm_io_read_pcd": (C.c_int, [C.c_char_p, _VP, C.c_int64, _VP]),
    "hgmm_p2p_export": (C.c_int, [_VP, _VP]),
    "hgmm_p2p_attach": (C.c_int, [_VP, _VP, C.c_int32]),
    "hgmm_p2p_detach": (C.c_int, [_VP]),
    "hgmm_p2p_enabled": (C.c_int, [_VP]),
    "hgmm_measure_fp32_peak": (C.c_int, [_VP, _VP]),
    "hgmm_sorted_points": (C.c_int, [_VP, _VP]),
    "hgmm_last_timing": (C.c_int, [_VP, _VP]),
    "hgmm_set_profiling": (C.c_int, [_VP, C.c_int]),
}


def load():
    """dlopen libhgmm.so and type every entry point.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HgmmError("libhgmm.so not built: run `python __graft_entry__.py build` (or make in %s)" % os.path.dirname(_HERE))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def f32c(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a
