"""CPU timings of the UNMODIFIED reference Python paths (VERDICT r1 item 5 / BASELINE.md section 3), run in the BUILD CONTAINER
(the reference checkout does not travel to the GPU box) -> profiles/r02_reference_python_cpu.json, which bench.py attaches to
its line as `reference_python_cpu` with this provenance.  The reference files are imported under oracle/refshim exactly as
oracle/make_golden.py does (cupy -> numpy facade, stubs for open3d / matplotlib); nothing of the reference is copied.

    python profiles/time_reference_python.py
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden as MG      # noqa: E402  (sets up the shim path and np.infty)
from oracle import hgmm_tree             # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return [{"lib": d.get("internal_api"), "threads": d.get("num_threads")} for d in threadpool_info()]
    except Exception:
        return None


def main():
    bun0 = np.load(os.path.join(ROOT, "tests", "golden", "bun000_xyz.npy"))
    out = {"where": "build container (no GPU), NOT the benchmark host", "cpu_count": os.cpu_count(), "blas": blas_threads(),
           "numpy": np.__version__, "reference_commit": "ff2ea430", "note": "wall clock (time.perf_counter) around the reference's own functions, "
           "unmodified, imported under oracle/refshim; no extrapolation"}
    # (1) flat diag fit, src/python/gmm_waymo/src/gmm_impl.py::train_gmm (:118-145)
    ref = MG.load_ref_flat()
    X = bun0.astype(np.float32)
    for J, its in ((8, 10), (800, 5)):
        rng = np.random.default_rng(0)
        m0 = X[rng.choice(len(X), J, replace=False)].astype(np.float32)
        c0 = (0.1 * np.ones((J, 3))).astype(np.float32)
        w0 = (np.ones(J) / J).astype(np.float32)
        quiet(ref.train_gmm, X, 1, 0.0, m0, c0, w0, "diag")
        t0 = time.perf_counter()
        quiet(ref.train_gmm, X, its, 0.0, m0, c0, w0, "diag")
        dt = time.perf_counter() - t0
        out["train_gmm_diag_J%d_bun000" % J] = {"function": "gmm_waymo/src/gmm_impl.py::train_gmm (cov_type='diag', xp=numpy, float32)",
                                                "points": int(len(X)), "components": J, "iterations": its, "seconds": dt,
                                                "ms_per_em_iteration": dt / its * 1e3, "mpoint_iters_per_sec": len(X) * its / dt / 1e6}
        print(J, out["train_gmm_diag_J%d_bun000" % J], flush=True)
    # (2) tree build, src/python/hgmm/hgmm_cupy_cpu_working.py::buildGMMTree (:122-160), pure-Python per-point loops
    ns = MG.load_ref_hgmm_cpu()
    Xs = bun0[::10][:4000].astype(np.float64)
    t0 = time.perf_counter()
    nodes = quiet(ns["buildGMMTree"], Xs, 2, 80.0, 1.0e-4)
    dt = time.perf_counter() - t0
    idx = hgmm_tree.reference_init_indices(2, flavor="cpu")
    _, _, _, _, oit, _ = hgmm_tree.build_gmm_tree(Xs, 2, 80.0, 1.0e-4, Xs[idx], sig2=0.00034, ll_mode="level", return_trace=True)
    out["buildGMMTree_4k_L2"] = {"function": "hgmm/hgmm_cupy_cpu_working.py::buildGMMTree (ls=80, ld=1e-4; its own seeded init)",
                                 "points": int(len(Xs)), "levels": 2, "seconds": dt, "em_iterations_oracle_same_init": [int(v) for v in oit],
                                 "seconds_per_point_iteration": dt / (len(Xs) * max(sum(oit), 1))}
    print(out["buildGMMTree_4k_L2"], flush=True)
    # (3) registration E-step + M-step on 2k target points, hgmm_gpu.py:550-577,729-752 (CPU twin in the same file)
    gt = ns["GMMTree"](None, tree_level=2, lambda_c=0.01)
    gt._source = Xs
    gt._nodes = nodes
    th = np.deg2rad(8.0)
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    T = (Xs @ Rz.T + np.array([0.004, -0.003, 0.002]))[::2]
    t0 = time.perf_counter()
    est = quiet(gt.expectation_step, T)
    t1 = time.perf_counter()
    quiet(gt.maximization_step, est, gt._tf_result)
    t2 = time.perf_counter()
    out["registration_step_2k"] = {"function": "gmmTreeRegESTep + GMMTree.maximization_step (hgmm_cupy_cpu_working.py twin of hgmm_gpu.py:550-577,729-752)",
                                   "target_points": int(len(T)), "tree_nodes": 72, "estep_seconds": t1 - t0, "mstep_seconds": t2 - t1,
                                   "iterations_per_sec": 1.0 / (t2 - t0)}
    print(out["registration_step_2k"], flush=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_reference_python_cpu.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
