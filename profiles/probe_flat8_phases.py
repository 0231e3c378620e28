"""Where a warp of the large-mixture sweep spends its cycles: per-phase clock64() counters of em_flat8_kernel on configs[1].
Needs the profiling build (`make prof` in the package -> build/libhgmm_prof.so); not a bench line.
usage: HGMM_LIB_PATH=<pkg>/build/libhgmm_prof.so python profiles/probe_flat8_phases.py > profiles/r02_flat8_phases.txt"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
os.environ.setdefault("HGMM_LIB_PATH", os.path.join(PKG, "build", "libhgmm_prof.so"))
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
import numpy as np, torch, hgmm_b200
from hgmm_b200 import _lib

lib = ctypes.CDLL(_lib.LIB_PATH)
eng = hgmm_b200.Engine(0)
X = np.load(os.path.join(ROOT, "tests/golden/bun000_xyz.npy"))
names = ["barrier wait", "finish + psi", "density pass", "moment pass", "prologue", "epilogue"]
for J in (800,):
    mu0 = X[np.random.default_rng(1).choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1)); w0 = np.full(J, 1 / J, np.float32)
    eng.set_points(torch.from_numpy(X).cuda())
    for tile, name in ((9, "em_flat8"), (11, "em_flat8 staggered"), (10, "em_flat8 cholesky-form")):
        for _ in range(3):
            eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, want_outputs=False, tile_points=tile)
        ms = eng.last_timing_ms()[0]
        buf = (ctypes.c_ulonglong * (148 * 16 * 8))()
        rc = lib.hgmm_debug_flat8_prof(buf)
        a = np.ctypeslib.as_array(buf).reshape(148, 16, 8)[:, :, :6].astype(np.float64)
        tot = a.sum(axis=2)
        print("J=%d %s: 10-iteration fit %.4f ms (rc %d); cycles per warp of the LAST sweep, mean over the 148 CTAs" % (J, name, ms, rc))
        print("  warp   " + "  ".join("%13s" % n for n in names) + "          total")
        for w in range(16):
            print("  %4d   " % w + "  ".join("%13.0f" % a[:, w, i].mean() for i in range(6)) + "  %13.0f" % tot[:, w].mean())
        heavy = a[:, :12, :].mean(axis=(0, 1))
        print("  heavy warps (0-11), share of their time: " + ", ".join("%s %.1f %%" % (names[i], 100 * heavy[i] / heavy.sum()) for i in range(6)))
        print("  slowest / fastest CTA (max over warps of the total): %.0f / %.0f cycles" % (tot.max(axis=1).max(), tot.max(axis=1).min()))
