"""hgmm_b200 -- Python face of libhgmm (B200 / sm_100a hierarchical-GMM fit + registration engine).

Thin ctypes wrapper that keeps the reference's Python surface:

    gmm_impl : train_gmm, predict, init_gmm_params, timer          (src/python/gmm_waymo/src/gmm_impl.py)
    gmm      : Feature, GMM_GPU, GMM_GPU_Base, GMM_CPU, GMM_CPU_Base (src/python/gmm_waymo/src/gmm.py)
    hgmm     : buildGMMTree, GMMTree, registration_gmmtree, RigidTransformation, EstepResult, MstepResult
                                                                     (src/python/hgmm/hgmm_gpu.py)
    gmmreg   : registration_gmmreg, RigidGMMReg, L2DistRegistration, RigidCostFunction, GMM_GPU (older diag fitter)
                                                                     (src/python/gmmreg_gpu/{gmmreg,cost_functions,gmm,gmm_impl}.py)
    stream   : StreamingSegmenter -- re-fit every k frames, hard-assign every frame      (src/python/gmm_waymo/src/run_gmm_waymo_gpu.py:39-50)
    io       : read_ply, read_pcd -- the viewers' / drivers' cloud readers              (src/c++/main.cpp:45-79, main_reg.cpp:106-161)
    dist     : point sharding + NCCL communicator bootstrap over torch.distributed
    engine   : Engine, the object wrapper over the C ABI (include/hgmm.h)

All compute runs in the CUDA library; there is no CPU fallback and nothing here imports `oracle`.
"""
from ._lib import HgmmError, LIB_PATH  # noqa: F401
from .engine import Engine  # noqa: F401
from . import gmm_impl, gmm, hgmm, gmmreg, dist, stream, io  # noqa: F401

__all__ = ["Engine", "HgmmError", "gmm_impl", "gmm", "hgmm", "gmmreg", "dist", "stream", "io", "LIB_PATH"]
