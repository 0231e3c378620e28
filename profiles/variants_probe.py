"""GPU probe (not a bench line): device time of one 10-iteration fit of configs[1] for every kernel variant."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200"))
import numpy as np, torch, hgmm_b200
X = np.load(os.path.join(ROOT, "tests/golden/bun000_xyz.npy"))
out = {}
eng = hgmm_b200.Engine(0)
eng.set_points(torch.from_numpy(X).cuda())
for J in (800, 1024, 640, 320):
    if len(sys.argv) > 1 and J != int(sys.argv[1]):
        continue
    rng = np.random.default_rng(1)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1)); w0 = np.full(J, 1 / J, np.float32)
    for name, variant, tile in (("v3_big_pb8", 0, 1), ("v5_staged", 0, 6), ("v7_pipelined", 0, 8)):
        eng.set_profiling(False)
        for _ in range(3):
            eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, want_outputs=False, variant=variant, tile_points=tile)
        tot = [float(eng.last_timing_ms()[0]) for _ in range(5) if eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, want_outputs=False, variant=variant, tile_points=tile)]
        eng.set_profiling(True)
        eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, want_outputs=False, variant=variant, tile_points=tile)
        k = eng.last_timing_ms()
        out["J%d_%s" % (J, name)] = {"fit10_ms_min": min(tot), "sweep_kernel_us": 1e3 * k[1] / max(k[2], 1)}
out["fp32_peak_tflops_imm_reg_packed"] = eng.measure_fp32_peak()
print(json.dumps(out, indent=1))
