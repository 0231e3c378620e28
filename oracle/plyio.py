"""Minimal ASCII-PLY vertex reader for the reference's data/*.ply fixtures.  TEST INFRASTRUCTURE ONLY.

Follows the header (`element vertex N`, `end_header`) instead of the reference viewer's
"skip 24 lines, stop at the first non-3-token line" heuristic (src/c++/main.cpp:45-79) and
does NOT apply the viewer's hard-coded similarity transforms (main.cpp:30-38) -- SURVEY.md 8d.
"""
import numpy as np


def read_ply_vertices(path):
    with open(path, "r") as f:
        n = None
        for line in f:
            tok = line.split()
            if len(tok) == 3 and tok[0] == "element" and tok[1] == "vertex":
                n = int(tok[2])
            if tok and tok[0] == "end_header":
                break
        if n is None:
            raise ValueError("no 'element vertex' in %s" % path)
        out = np.empty((n, 3), dtype=np.float32)
        for i in range(n):
            out[i] = [float(v) for v in f.readline().split()[:3]]
    return out
