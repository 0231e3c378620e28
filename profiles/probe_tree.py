"""Probe (not a bench line): time the tree build of configs[2] / configs[4] on one GPU under the current environment
switches and print one JSON line per run.    python profiles/probe_tree.py <n> <L> <seed> <mode> [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200"))
import numpy as np, torch, hgmm_b200
from hgmm_b200 import hgmm as H, synth

n, L, seed, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cache = "/tmp/lidar_%d_%d.npy" % (n, seed)
if os.path.exists(cache):
    P = np.load(cache)
else:
    P = synth.lidar_sweep(n, seed=seed)
    np.save(cache, P)
init = P[H.reference_init_indices(L)]
eng = hgmm_b200.Engine(0)
eng.set_points(torch.from_numpy(P).cuda())
best = None
for r in range(reps):
    t0 = time.perf_counter()
    res = eng.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4.0, ll_mode=mode, want_current=False, want_outputs=(r == reps - 1))
    wall = (time.perf_counter() - t0) * 1e3
    ms = float(eng.last_timing_ms()[0])
    if best is None or ms < best:
        best = ms
its = int(res["iters"].sum())
out = {"n": n, "L": L, "mode": mode, "build_ms": best, "wall_ms_last": wall, "iters": res["iters"].tolist(), "em_iterations": its,
       "us_per_iteration": best * 1e3 / max(its, 1), "env": {k: v for k, v in os.environ.items() if k.startswith("HGMM_")},
       "pi_sum_leaf": float(res["pi"][H.level(L - 1):].sum()), "finite": bool(np.isfinite(res["mu"]).all() and np.isfinite(res["cov"]).all())}
print("PROBE " + json.dumps(out), flush=True)
