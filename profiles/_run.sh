cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sorted_sweep or (staged and (9 or 10 or 11 or 12))" 2>&1 | tail -15 | cut -c1-900) > gpurun_out/r2_flat8_tests.log
cat gpurun_out/r2_flat8_tests.log
cat > /tmp/ab.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
import hgmm_b200
eng = hgmm_b200.Engine(0)
X = np.load("tests/golden/bun000_xyz.npy")
rf = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64)))
for J in (800, 1024):
    mu0 = X[np.random.default_rng(1).choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1)); w0 = np.full(J, 1 / J, np.float32)
    eng.set_points(torch.from_numpy(X).cuda())
    ref = None
    for tile, name in ((8, "em_flat7"), (9, "em_flat8"), (10, "em_flat8_chol"), (12, "em_flat8_chol_stagger")):
        r = eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, tile_points=tile)
        if ref is None: ref = r
        d = max(rf(r["means"], ref["means"]), rf(r["covs"], ref["covs"]), rf(r["weights"], ref["weights"]))
        eng.set_profiling(True)
        best, k = 1e9, 1e9
        for _ in range(6):
            eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, want_outputs=False, tile_points=tile)
            tm = eng.last_timing_ms(); k = min(k, tm[1] / max(tm[2], 1))
        eng.set_profiling(False)
        for _ in range(6):
            eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, want_outputs=False, tile_points=tile)
            best = min(best, eng.last_timing_ms()[0])
        print("SWEEPAB sleep=%s J=%d %s: sweep %.2f us/launch, 10-iteration fit %.4f ms, vs em_flat7 %.3e" % (os.environ.get("HGMM_FLAT8_SLEEP_NS", "-"), J, name, k * 1e3, best, d), flush=True)
PY
for S in 0 1000 2000 3000; do HGMM_FLAT8_SLEEP_NS=$S python /tmp/ab.py 2>&1 | grep SWEEPAB; done
HGMM_FLAT8_SLEEP_NS=2000 timeout 300 python profiles/probe_flat8_phases.py > gpurun_out/r2_flat8_phases.txt 2>&1
grep -A3 "share of their time" gpurun_out/r2_flat8_phases.txt | cut -c1-250
