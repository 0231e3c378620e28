cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sorted_sweep" 2>&1 | tail -15 | cut -c1-600) > gpurun_out/r2_flat8_tests.log
cat gpurun_out/r2_flat8_tests.log
python - <<'PY' 2>&1 | tail -20
import os, sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
import hgmm_b200
eng = hgmm_b200.Engine(0)
X = np.load("tests/golden/bun000_xyz.npy")
rf = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64)))
for J in (800, 1024, 640):
    mu0 = X[np.random.default_rng(1).choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1)); w0 = np.full(J, 1 / J, np.float32)
    eng.set_points(torch.from_numpy(X).cuda())
    ref = None
    for tile, name in ((8, "em_flat7"), (9, "em_flat8"), (11, "em_flat8_stagger"), (10, "em_flat8_chol"), (12, "em_flat8_chol_stagger")):
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
        print("SWEEPAB J=%d %s: sweep %.2f us/launch, 10-iteration fit %.4f ms, vs em_flat7 %.1e" % (J, name, k * 1e3, best, d), flush=True)
PY
python - <<'PY' > /dev/null
import json
for f in ():
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f ms/step %.4f e2e %.1f sweep %.2f us frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_us"], d["roofline_fp32"]["frac"]))
    except Exception as e:
        print(f, "ERR", e)
PY
