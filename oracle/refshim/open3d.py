"""stub of the open3d names the reference touches at import / class-definition time
(hgmm/hgmm_gpu.py:588-589, hgmm_cupy_cpu_working.py:239)."""
__version__ = "0.9.0.0"


class _Vec(list):
    pass


class utility:  # noqa: N801
    Vector3dVector = _Vec


class geometry:  # noqa: N801
    class PointCloud:
        pass
