"""Run the reference's own CUDA fitter (oracle/_ref/ref_gmm_cuda, built by build_ref.sh) on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  For each J: (1) the reference's own cudaEvent time of its 10-iteration
loop (gmm_kernels.cu:483-488) -- the "reference CUDA" comparator of BASELINE.json configs[1]; (2) its fitted
means/weights against oracle.flat_gmm.cpp_fit(sigma_bug=True) started from the same rand() draw -- this pins the
oracle's restatement of the C++ variant (including the Sigma-for-Sigma^-1 bug, :97-103) to the reference itself.
Writes a JSON summary to stdout.
"""
import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import flat_gmm  # noqa: E402


def glibc_rand_indices(count, n):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)                      # the process default; the reference never calls srand (gmm_kernels.cu:375)
    return np.array([libc.rand() % n for _ in range(count)])


def main():
    exe = os.path.join(HERE, "_ref", "ref_gmm_cuda")
    pts = os.path.join(ROOT, "tests", "golden", "bun000_xyz.npy")
    X = np.load(pts)
    out = {}
    for J in (8, 100, 800):
        with tempfile.TemporaryDirectory() as td:
            ob = os.path.join(td, "out.bin")
            r = subprocess.run([exe, pts, str(J), "10", ob], capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                out["J%d" % J] = {"error": r.stderr[-300:]}
                continue
            m = re.search(r"Time elapsed: ([0-9.]+)", r.stdout)
            raw = np.fromfile(ob, dtype=np.float32)
        mu, w = raw[:3 * J].reshape(J, 3), raw[3 * J:]
        rec = {"ref_cuda_ms_10_iterations": float(m.group(1)) if m else None,
               "ref_cuda_em_iters_per_sec": 10.0 / (float(m.group(1)) * 1e-3) if m else None,
               "finite": bool(np.isfinite(mu).all() and np.isfinite(w).all())}
        if J <= 100:
            idx = glibc_rand_indices(J, len(X))
            ow, omu, ocov, _ = flat_gmm.cpp_fit(X, X[idx], 10, sigma0_sq=1.0, sigma_bug=True)
            rec["oracle_sigma_bug_vs_ref_cuda"] = {"means": flat_gmm.rel_fro(omu, mu), "weights": flat_gmm.rel_fro(ow, w)}
            ow2, omu2, _, _ = flat_gmm.cpp_fit(X, X[idx], 10, sigma0_sq=1.0, sigma_bug=False)
            rec["oracle_fixed_metric_vs_ref_cuda"] = {"means": flat_gmm.rel_fro(omu2, mu), "weights": flat_gmm.rel_fro(ow2, w)}
        out["J%d" % J] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
