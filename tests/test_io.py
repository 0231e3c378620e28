"""CPU: the ingestion readers (hgmm_io_read_ply / hgmm_io_read_pcd, include/hgmm.h) -- SURVEY.md 8f-1.  Fixtures are written
by the test with the layout of the reference's data files; the real files are used as well when the reference checkout is
present (never on the GPU box)."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLD

REF = "/root/reference"

BUNNY_HEADER = """ply
format ascii 1.0
obj_info is_cyberware_data 1
obj_info is_mesh 0
obj_info is_warped 0
obj_info is_interlaced 1
obj_info num_cols 512
obj_info num_rows 400
obj_info echo_rgb_offset_x 0.013000
obj_info echo_rgb_offset_y 0.153600
obj_info echo_rgb_offset_z 0.172000
obj_info echo_rgb_frontfocus 0.930000
obj_info echo_rgb_backfocus 0.012660
obj_info echo_rgb_pixelsize 0.000010
obj_info echo_rgb_centerpixel 232
obj_info echo_frames 512
obj_info echo_lgincr 0.000500
element vertex %d
property float x
property float y
property float z
element range_grid %d
property list uchar int vertex_indices
end_header
"""


def write_bunny_like(path, pts, eol="\n"):
    """the Stanford range-scan layout of data/bun000.ply: 24 header lines (vertex count on line 18), vertex lines with a
    trailing blank, then range_grid lines of one or two tokens"""
    grid = ["0", "1 0", "1 1", "0"]
    txt = BUNNY_HEADER % (len(pts), len(grid))
    txt += "".join("%.9g %.9g %.9g \n" % tuple(p) for p in pts)
    txt += "".join(g + " \n" for g in grid)
    with open(path, "w", newline="") as f:
        f.write(txt.replace("\n", eol))


@pytest.mark.parametrize("mode", ["header", "viewer_fit", "viewer_reg"])
@pytest.mark.parametrize("eol", ["\n", "\r\n"])
def test_ply_readers_agree_on_the_bunny_layout(tmp_path, mode, eol):
    from hgmm_b200 import io
    pts = np.random.default_rng(0).normal(0, 0.05, (257, 3)).astype(np.float32)
    p = tmp_path / "scan.ply"
    write_bunny_like(p, pts, eol)
    got = io.read_ply(p, mode)
    assert got.dtype == np.float32 and got.shape == (257, 3)
    assert np.array_equal(got, pts)                          # %.9g round-trips float32 exactly


def test_ply_header_mode_handles_other_headers_and_extra_properties(tmp_path):
    from hgmm_b200 import io
    p = tmp_path / "small.ply"
    p.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 3\nproperty float x\nproperty float y\n"
                 "property float z\nproperty float confidence\nelement face 1\nproperty list uchar int vertex_indices\n"
                 "end_header\n1 2 3 0.5\n4 5 6 0.25\n-7 8.5 9e-1 1\n3 0 1 2\n")
    assert np.array_equal(io.read_ply(p), np.array([[1, 2, 3], [4, 5, 6], [-7, 8.5, 0.9]], np.float32))
    # viewer_fit counts lines instead of parsing: with this 11-line header it would start in the wrong place -> not 3 points
    assert len(io.read_ply(p, "viewer_fit")) != 3


def test_ply_errors(tmp_path):
    from hgmm_b200 import io, HgmmError
    with pytest.raises(HgmmError):
        io.read_ply(tmp_path / "missing.ply")
    b = tmp_path / "binary.ply"
    b.write_text("ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
    with pytest.raises(HgmmError):
        io.read_ply(b)
    t = tmp_path / "truncated.ply"
    t.write_text("ply\nformat ascii 1.0\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\nend_header\n1 2 3\n")
    with pytest.raises(HgmmError):
        io.read_ply(t)
    e = tmp_path / "empty.ply"
    e.write_text("ply\nformat ascii 1.0\nelement vertex 0\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
    assert io.read_ply(e).shape == (0, 3)


PCD_HEADER = "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH %d\nHEIGHT 1\n" \
             "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n"


def test_pcd_binary_and_ascii(tmp_path):
    from hgmm_b200 import io, HgmmError
    pts = np.random.default_rng(1).normal(0, 20, (1001, 3)).astype(np.float32)
    b = tmp_path / "sweep.pcd"                            # the layout of src/python/hgmm/waymo*.pcd
    with open(b, "wb") as f:
        f.write((PCD_HEADER % (len(pts), len(pts), "binary")).encode())
        f.write(pts.tobytes())
    assert np.array_equal(io.read_pcd(b), pts)            # bit-exact
    a = tmp_path / "ascii.pcd"
    with open(a, "w") as f:
        f.write(PCD_HEADER % (len(pts), len(pts), "ascii"))
        f.write("".join("%.9g %.9g %.9g\n" % tuple(p) for p in pts))
    assert np.array_equal(io.read_pcd(a), pts)
    x = tmp_path / "xyzi.pcd"                             # a fourth field: rows are 16 bytes, xyz lead
    rows = np.concatenate([pts, np.ones((len(pts), 1), np.float32)], axis=1)
    with open(x, "wb") as f:
        f.write(("VERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nPOINTS %d\nDATA binary\n"
                 % (len(pts), len(pts))).encode())
        f.write(rows.tobytes())
    assert np.array_equal(io.read_pcd(x), pts)
    c = tmp_path / "lzf.pcd"
    c.write_text(PCD_HEADER % (4, 4, "binary_compressed"))
    with pytest.raises(HgmmError):
        io.read_pcd(c)
    s = tmp_path / "short.pcd"
    with open(s, "wb") as f:
        f.write((PCD_HEADER % (10, 10, "binary")).encode())
        f.write(struct.pack("<6f", *range(6)))
    with pytest.raises(HgmmError):
        io.read_pcd(s)


@pytest.mark.parametrize("header", [
    "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1\n",          # COUNT shorter than FIELDS
    "FIELDS x y z intensity\nSIZE 4 4 4 -4\nTYPE F F F F\nCOUNT 1 1 1 1\n",       # negative SIZE on a trailing field
    "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 -7\n",       # negative COUNT
    "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 0\n",        # zero COUNT
    "FIELDS x y z blob\nSIZE 4 4 4 8\nTYPE F F F F\nCOUNT 1 1 1 2000000000\n",    # absurd row size
])
def test_pcd_hostile_headers_are_rejected_not_fatal(tmp_path, header):
    """untrusted headers: a bad COUNT / SIZE must come back as HGMM_ERR_IO, never read out of bounds, never let an
    exception cross the C ABI"""
    from hgmm_b200 import io, HgmmError
    f = tmp_path / "bad.pcd"
    with open(f, "wb") as fh:
        fh.write(("VERSION 0.7\n" + header + "WIDTH 4\nHEIGHT 1\nPOINTS 4\nDATA binary\n").encode())
        fh.write(b"\0" * 256)
    with pytest.raises(HgmmError):
        io.read_pcd(f)


def test_c_abi_two_pass_protocol(tmp_path):
    """capacity smaller than the file: the count is still reported, only `capacity` points are written"""
    import ctypes as C
    from hgmm_b200 import _lib
    lib = _lib.load()
    pts = np.arange(30, dtype=np.float32).reshape(10, 3)
    p = tmp_path / "ten.ply"
    write_bunny_like(p, pts)
    n = C.c_int64(0)
    assert lib.hgmm_io_read_ply(str(p).encode(), 0, None, 0, C.byref(n)) == 0 and n.value == 10
    buf = np.full((4, 3), -1, np.float32)
    assert lib.hgmm_io_read_ply(str(p).encode(), 2, buf.ctypes.data_as(C.c_void_p), 4, C.byref(n)) == 0 and n.value == 10
    assert np.array_equal(buf, pts[:4])
    assert lib.hgmm_io_read_ply(None, 0, None, 0, C.byref(n)) == -1
    assert lib.hgmm_io_read_ply(str(p).encode(), 7, None, 0, C.byref(n)) == -1
    assert lib.hgmm_io_read_ply(str(tmp_path / "nope.ply").encode(), 0, None, 0, C.byref(n)) == -6


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_real_reference_files():
    from hgmm_b200 import io
    want = np.load(os.path.join(GOLD, "bun000_xyz.npy"))
    for mode in ("header", "viewer_fit", "viewer_reg"):
        got = io.read_ply(os.path.join(REF, "data", "bun000.ply"), mode)
        assert got.shape == (40256, 3) and np.array_equal(got, want), mode
    assert io.read_ply(os.path.join(REF, "data", "bun045.ply")).shape == (40097, 3)
    w = io.read_pcd(os.path.join(REF, "src", "python", "hgmm", "waymo1.pcd"))
    assert w.shape == (14637, 3) and np.isfinite(w).all() and np.abs(w).max() < 80.0
