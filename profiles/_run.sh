cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2_all.log
grep -E "passed|failed|FAILED" gpurun_out/r2_all.log | head -20
rm -f gpurun_out/r2_probe5.log
for cfg in "100000 4 2024 estep" "1000000 5 2025 estep"; do
  timeout 300 python profiles/probe_tree.py $cfg 4 2>&1 | grep -E "PROBE|PROF|Error|error" >> gpurun_out/r2_probe5.log
  HGMM_TREE_PROF=1 timeout 300 python profiles/probe_tree.py $cfg 2 2>&1 | grep -E "PROF|Error|error" | tail -1 >> gpurun_out/r2_probe5.log
done
cat gpurun_out/r2_probe5.log
python - <<'PY' 2>&1 | tail -6
import os, sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
import hgmm_b200
from hgmm_b200 import hgmm as H
S = np.load("tests/golden/bun000_xyz.npy"); T = np.load("tests/golden/bun045_xyz.npy")
eng = hgmm_b200.Engine(0)
for L in (3, 4):
    init = S[H.reference_init_indices(L)]
    eng.set_points(torch.from_numpy(S).cuda()); eng.reg_set_target(torch.from_numpy(T).cuda())
    eng.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep", want_current=False, want_outputs=False)
    for solver in ("twist_lstsq", "procrustes_svd"):
        best = 1e9
        for _ in range(5):
            rot, t, q, it, _h = eng.register_tree(solver=solver, maxiter=20, tol=0.0)
            best = min(best, float(eng.last_timing_ms()[0]))
        print("REG L=%d %s: %d iterations %.3f ms -> %.1f us/iteration  q=%.9g" % (L, solver, it, best, best * 1e3 / it, q))
PY
