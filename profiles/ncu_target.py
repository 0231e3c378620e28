"""ncu target: a few fits of configs[1] (J=800 full-cov on bun000) and of the 100k LiDAR tree; run under
`ncu -k regex:<kernel>`; not a bench line."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200"))
import numpy as np, torch, hgmm_b200
what = sys.argv[1] if len(sys.argv) > 1 else "flat"
eng = hgmm_b200.Engine(0)
if what == "flat":
    X = np.load(os.path.join(ROOT, "tests/golden/bun000_xyz.npy"))
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    tile = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rng = np.random.default_rng(1)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1)); w0 = np.full(J, 1 / J, np.float32)
    eng.set_points(torch.from_numpy(X).cuda())
    for _ in range(3):
        eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, want_outputs=False, tile_points=tile)
elif what == "reg":
    from hgmm_b200 import hgmm as H
    S = np.load(os.path.join(ROOT, "tests/golden/bun000_xyz.npy")); T = np.load(os.path.join(ROOT, "tests/golden/bun045_xyz.npy"))
    init = S[H.reference_init_indices(3)]
    eng.set_points(torch.from_numpy(S).cuda()); eng.reg_set_target(torch.from_numpy(T).cuda())
    for _ in range(2):
        eng.fit_tree(init, 3, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep", want_current=False, want_outputs=False)
        eng.register_tree(solver="twist_lstsq", maxiter=20, tol=1e-4)
else:
    from hgmm_b200 import hgmm as H, synth
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    P = synth.lidar_sweep(n, seed=2024)
    init = P[H.reference_init_indices(L)]
    eng.set_points(torch.from_numpy(P).cuda())
    for _ in range(2):
        eng.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4.0, ll_mode="estep", want_current=False, want_outputs=False)
