"""Cloud ingestion through libhgmm's host-side readers (include/hgmm.h: hgmm_io_read_ply / hgmm_io_read_pcd) -- the step the
reference's viewers and drivers run in front of the fit (src/c++/main.cpp:45-79, src/c++/main_reg.cpp:106-161, the Open3D
`read_point_cloud` calls of src/python/hgmm/hgmm_gpu.py:813-822).  Returns [N,3] float32 arrays, the layout
Engine.set_points takes.  No GPU is needed for these calls."""
import ctypes as C

import numpy as np

from . import _lib as L

PLY_MODES = {"header": 0, "viewer_fit": 1, "viewer_reg": 2}


def _read(call, what, path):
    n = C.c_int64(0)
    rc = call(None, 0, C.byref(n))
    if rc != L.HGMM_OK:
        raise L.HgmmError("%s(%r) failed (status %d): cannot open the file or unsupported format" % (what, path, rc))
    out = np.empty((int(n.value), 3), np.float32)
    if n.value:
        rc = call(out.ctypes.data_as(C.c_void_p), int(n.value), C.byref(n))
        if rc != L.HGMM_OK or int(n.value) != len(out):
            raise L.HgmmError("%s(%r) failed on the second pass (status %d)" % (what, path, rc))
    return out


def read_ply(path, mode="header"):
    """ASCII PLY vertices.  mode: 'header' (header-driven), 'viewer_fit' (main.cpp's reader: 24 header lines, then lines
    while they have three tokens) or 'viewer_reg' (main_reg.cpp's reader: vertex count on line 18, 6 more header lines)."""
    lib = L.load()
    m = PLY_MODES[mode]
    p = str(path).encode()
    return _read(lambda buf, cap, n: lib.hgmm_io_read_ply(p, m, buf, cap, n), "hgmm_io_read_ply", path)


def read_pcd(path):
    """PCD v0.7 with x y z float32 leading every point, DATA ascii or binary."""
    lib = L.load()
    p = str(path).encode()
    return _read(lambda buf, cap, n: lib.hgmm_io_read_pcd(p, buf, cap, n), "hgmm_io_read_pcd", path)
