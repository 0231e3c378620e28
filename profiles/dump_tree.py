"""debug helper: run config-size tree builds on the GPU and save the models under gpurun_out/ for offline comparison"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200"))
import numpy as np, hgmm_b200
from hgmm_b200 import hgmm as H, synth
tag = sys.argv[1]
eng = hgmm_b200.Engine(0)
for name, L, fixed in (("lidar100k_L4", 4, 2), ("lidar50k_L5", 5, 2), ("lidar100k_L4", 4, 12), ("lidar50k_L5", 5, 10)):
    if name == "lidar100k_L4":
        P = synth.lidar_sweep(100000, seed=2024)
    else:
        P = synth.lidar_sweep(1000000, seed=2025)
        P = P[np.random.default_rng(0).permutation(len(P))[:50000]]
    init = P[H.reference_init_indices(L)]
    eng.set_points(P)
    r = eng.fit_tree(init, L, ls=0.0, ld=1e-4, sig2=4.0, ll_mode="estep", max_iters_per_level=fixed)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_%s_%s_fixed%d.npz" % (tag, name, fixed)), pi=r["pi"], mu=r["mu"], cov=r["cov"],
                        iters=r["iters"], q=r["q"], current=r["current"].astype(np.int32))
    print(tag, name, fixed, r["iters"].tolist(), r["q"].tolist(), flush=True)
