"""GPU: the CUDA path (through the C ABI) against the oracle and the committed golden vectors.

Tolerance (BASELINE.json north_star): fitted mu / Sigma / pi and the recovered SE(3) within 1e-4
relative Frobenius of the reference semantics (float64 oracle, pinned to the unmodified reference by
oracle/make_golden.py).  Where the oracle is slow, size-independent properties are checked instead."""
import os

import numpy as np
import pytest

from conftest import gold, rel_fro

pytestmark = pytest.mark.gpu
TOL = 1e-4


# ------------------------------------------------------------------ flat, python variant (C1)
@pytest.mark.parametrize("cov_type", ["diag", "spherical"])
@pytest.mark.parametrize("tag", ["sub4k_J8", "bun000_J8", "sub4k_J32"])
def test_flat_py_matches_reference_golden(engine, bun000, cov_type, tag):
    g = gold("flat_py_%s_%s.npz" % (cov_type, tag))
    X = bun000[::int(g["stride"])]
    engine.set_points(X)
    r = engine.fit_flat(g["means0"], g["covs0"], g["weights0"], cov_type=cov_type, max_iter=10, tol=0.0)
    assert r["iters"] == 10
    assert rel_fro(r["means"], g["ref_means"]) < TOL
    assert rel_fro(r["covs"], g["ref_covs"]) < TOL
    assert rel_fro(r["weights"], g["ref_weights"]) < TOL
    assert rel_fro(r["inv_cov"], g["ref_inv_cov"]) < TOL
    assert rel_fro(r["ll"], g["ref_ll"]) < TOL
    lab = engine.predict_flat()
    assert (lab == g["ref_labels"]).mean() > 0.999


def test_train_gmm_api_is_drop_in(engine, bun000):
    from hgmm_b200 import gmm_impl
    g = gold("flat_py_diag_sub4k_J8.npz")
    X = bun000[::10]
    inv, mu, w, cov, ll = gmm_impl.train_gmm(X, 10, 0.0, g["means0"], g["covs0"], g["weights0"], "diag", engine=engine)
    assert rel_fro(mu, g["ref_means"]) < TOL and rel_fro(cov, g["ref_covs"]) < TOL and len(ll) == 10
    lab = gmm_impl.predict(X, inv, mu, w, "diag", engine=engine)
    assert (lab == g["ref_labels"]).mean() > 0.999


def test_flat_py_tolerance_stops_early(engine, bun000):
    from oracle import flat_gmm
    g = gold("flat_py_diag_sub4k_J8.npz")
    X = bun000[::10]
    engine.set_points(X)
    r = engine.fit_flat(g["means0"], g["covs0"], g["weights0"], cov_type="diag", max_iter=60, tol=1e-3)
    o = flat_gmm.py_train_gmm(X, 60, 1e-3, g["means0"], g["covs0"], g["weights0"], "diag")
    assert r["iters"] == len(o[4]) < 60
    assert rel_fro(r["means"], o[1]) < TOL and rel_fro(r["covs"], o[3]) < TOL


# ------------------------------------------------------------------ flat, C++ variant (C2)
@pytest.mark.parametrize("J,stride,sig", [(8, 10, 4e-4), (100, 10, 1e-4), (33, 7, 2e-4), (256, 2, 1e-4)])
def test_flat_full_matches_oracle(engine, bun000, J, stride, sig):
    from oracle import flat_gmm
    X = bun000[::stride]
    rng = np.random.default_rng(1)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * np.float32(sig), (J, 1, 1))
    w0 = np.full(J, 1.0 / J, np.float32)
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10)
    ow, omu, ocov, oll = flat_gmm.cpp_fit(X, mu0, 10, sigma0_sq=np.float32(sig))
    assert rel_fro(r["weights"], ow) < TOL
    assert rel_fro(r["means"], omu) < TOL
    assert rel_fro(r["covs"], ocov) < TOL
    assert rel_fro(r["ll"], oll) < TOL
    assert abs(float(r["weights"].sum()) - 1.0) < 1e-5


@pytest.mark.parametrize("J,sig", [(800, 1e-4), (800, 1.0), (1024, 1e-4)])
def test_flat_full_config2_matches_c_oracle(engine, bun000, J, sig):
    """config 2 at full size (40 256 pts x J=800, 10 iterations, both initial variances of SURVEY.md 8d) against the
    plain-C float64 oracle (itself checked against the NumPy oracle in test_c_oracle_agrees_with_numpy_oracle)"""
    from oracle import c_oracle
    X = bun000
    rng = np.random.default_rng(1)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * np.float32(sig), (J, 1, 1))
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, np.full(J, 1.0 / J, np.float32), cov_type="full", max_iter=10)
    ow, omu, ocov, oll = c_oracle.flat_fit(X, mu0, 10, np.float32(sig))
    assert np.isfinite(ocov).all()
    errs = (rel_fro(r["weights"], ow), rel_fro(r["means"], omu), rel_fro(r["covs"], ocov), rel_fro(r["ll"], oll))
    print("J=%d sig=%g rel_fro (w, mu, cov, ll) = %.2e %.2e %.2e %.2e" % ((J, sig) + errs))
    if sig == 1.0:
        # Sigma0 = I on a 0.15 m object: 800 near-identical components whose differences double every iteration
        # (tests/test_oracle_golden.py::test_identity_start_amplifies_rounding): fp32 parameter storage alone puts a
        # float64 EM 4e-5 off after 10 iterations, so 1e-4 is not attainable by ANY fp32 pipeline, the reference's included.
        assert errs[3] < TOL and errs[1] < 5 * TOL and errs[2] < 5 * TOL and errs[0] < 30 * TOL, errs
    else:
        assert max(errs) < TOL, errs


@pytest.mark.parametrize("variant,tile", [(0, 0), (0, 1), (0, 2), (0, 3), (0, 4), (0, 6), (0, 7), (0, 8), (3, 0), (2, 0), (2, 1), (1, 64), (1, 128), (1, 256), (1, 512)])
def test_flat_kernel_variants_agree(engine, bun000, variant, tile):
    from oracle import flat_gmm
    X = bun000[::5]
    rng = np.random.default_rng(2)
    J = 600 if tile in (3, 4, 6, 7, 8) else 96   # the two-team / pipelined builds need >= 17 component slots
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 2e-4, (J, 1, 1))
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=4, tile_points=tile, variant=variant)
    ow, omu, ocov, _ = flat_gmm.cpp_fit(X, mu0, 4, sigma0_sq=np.float32(2e-4))
    assert rel_fro(r["means"], omu) < TOL and rel_fro(r["covs"], ocov) < TOL and rel_fro(r["weights"], ow) < TOL


STAGED = [(11, 353), (11, 800), (12, 740), (12, 1024), (10, 320), (10, 353), (10, 740), (10, 800), (10, 1024), (9, 320), (9, 353), (9, 545), (9, 740), (9, 800), (9, 833), (9, 1024), (8, 320), (8, 353), (8, 385), (8, 480), (8, 545), (8, 600), (8, 700), (8, 740), (8, 800), (8, 833), (8, 897), (8, 1024), (6, 160), (6, 161), (6, 320), (6, 545), (6, 800), (6, 1024), (7, 225), (7, 256), (7, 320), (7, 545), (7, 800), (7, 1024)]


@pytest.mark.parametrize("tile,J", STAGED)
def test_flat_staged_kernel_matches_oracle(engine, bun000, tile, J):
    """flat_em5.cu / flat_em6.cu / flat_em7.cu / flat_em8.cu (densities staged in shared memory, tile_points = 6: component pair
    per lane, packed FP32, CTA barriers; 7: one component per thread; 8: mbarrier chunk pipeline; 9: the same with the moment
    pass about one origin per chunk over the cell-sorted cloud; 10: that with the Cholesky-form density
    pass; 11 / 12: 9 / 10 with every other warp of a scheduler taking the moment pass first): every warp count, ragged J and a cloud whose per-CTA share is not a multiple of the
    chunk or of the 8-point batch"""
    from oracle import flat_gmm
    X = bun000[::3][:13001]
    rng = np.random.default_rng(J)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 3e-4, (J, 1, 1))
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=3, tile_points=tile)
    ow, omu, ocov, oll = flat_gmm.cpp_fit(X, mu0, 3, sigma0_sq=np.float32(3e-4))
    assert rel_fro(r["means"], omu) < TOL and rel_fro(r["covs"], ocov) < TOL and rel_fro(r["weights"], ow) < TOL
    assert rel_fro(r["ll"], oll) < TOL
    r2 = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=3, tile_points=tile)
    assert np.array_equal(r["means"], r2["means"]) and np.array_equal(r["covs"], r2["covs"])       # bit-reproducible


@pytest.mark.parametrize("tile", [6, 7, 8, 9, 10])
def test_flat_staged_kernel_small_and_large_clouds(engine, bun000, tile):
    """fewer points than CTAs x 16 (short grid), and more than 512 points per CTA (several staging rounds)"""
    from oracle import flat_gmm
    J = 260
    for X in (bun000[::40], np.concatenate([bun000, bun000[::2] + np.float32(1e-4), bun000[::3] - np.float32(1e-4)])):
        rng = np.random.default_rng(7)
        mu0 = X[rng.choice(len(X), J, replace=False)]
        cov0 = np.tile(np.eye(3, dtype=np.float32) * 3e-4, (J, 1, 1))
        engine.set_points(X)
        r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=3, tile_points=tile)
        ow, omu, ocov, oll = flat_gmm.cpp_fit(X, mu0, 3, sigma0_sq=np.float32(3e-4))
        assert rel_fro(r["means"], omu) < TOL and rel_fro(r["covs"], ocov) < TOL and rel_fro(r["weights"], ow) < TOL
        assert rel_fro(r["ll"], oll) < TOL


@pytest.mark.parametrize("tile", [6, 7, 8, 9, 10])
@pytest.mark.parametrize("cov_type", ["diag", "spherical"])
def test_flat_staged_kernel_py_flavour(engine, bun000, cov_type, tile):
    """gmm_impl.py semantics (log(sum exp + 1e-8), +1e-6 floors) through the staged kernels, J = 260"""
    from oracle import flat_gmm
    X = bun000[::4]
    J = 260
    rng = np.random.default_rng(11)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.full((J, 3) if cov_type == "diag" else (J,), 1e-3, np.float32)
    w0 = np.full(J, 1 / J, np.float32)
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, w0, cov_type=cov_type, max_iter=5, tol=0.0, tile_points=tile)
    o = flat_gmm.py_train_gmm(X, 5, 0.0, mu0, cov0, w0, cov_type)
    assert rel_fro(r["means"], o[1]) < TOL and rel_fro(r["weights"], o[2]) < TOL and rel_fro(r["covs"], o[3]) < TOL
    assert rel_fro(r["ll"], o[4]) < TOL


@pytest.mark.parametrize("tile", [6, 7, 8, 9, 10])
def test_flat_staged_kernel_far_points(engine, tile):
    """the staged kernels' exact (max-shifted) path: 30-60 sigma outliers inside otherwise ordinary chunks"""
    from oracle import flat_gmm
    rng = np.random.default_rng(4)
    X = np.concatenate([rng.normal(0, 0.01, (3000, 3)), rng.normal(0, 0.01, (3000, 3)) + [1.0, 0, 0],
                        [[0.5, 0.3, 0.0], [0.45, -0.2, 0.1], [0.3, 0.3, 0.3]]]).astype(np.float32)
    J = 288
    mu0 = X[rng.choice(len(X) - 3, J, replace=False)]          # never an outlier
    X = X[rng.permutation(len(X))]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1))
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=2, tile_points=tile)
    ow, omu, ocov, oll = flat_gmm.cpp_fit(X, mu0, 2, sigma0_sq=np.float32(1e-4))
    assert rel_fro(r["weights"], ow) < TOL and rel_fro(r["means"], omu) < TOL
    assert rel_fro(r["covs"], ocov) < TOL and rel_fro(r["ll"], oll) < TOL


def _morton_cells(X):
    """csrc/cloud_sort.cu cell_of, operation for operation in float32"""
    lo, hi = X.min(axis=0), X.max(axis=0)
    ext = np.float32(max(float((hi - lo).max()), 1e-30))
    sc = np.float32(16.0) / ext
    t = ((X - lo).astype(np.float32) * sc).astype(np.float32)
    g = np.minimum(np.where(t > 0, np.minimum(t, np.float32(1e6)), 0).astype(np.int64), 15)
    spread = lambda v: (v & 1) | ((v & 2) << 2) | ((v & 4) << 4) | ((v & 8) << 6)
    return spread(g[:, 0]) | (spread(g[:, 1]) << 1) | (spread(g[:, 2]) << 2)


@pytest.mark.parametrize("n", [1, 255, 257, 40256, 300001])
def test_cloud_sort_is_a_stable_cell_sort(engine, bun000, n):
    """the order em_flat8_kernel reads the cloud in: a permutation of the input, cells non-decreasing, input order kept inside a
    cell (so a fit is a pure function of its input), block boundaries (256-point blocks, several points per thread) included"""
    rng = np.random.default_rng(n)
    X = bun000[:n] if n <= len(bun000) else (bun000[rng.integers(0, len(bun000), n)] + rng.normal(0, 1e-3, (n, 3))).astype(np.float32)
    engine.set_points(X)
    S = engine.sorted_points()
    cells = _morton_cells(X)
    order = np.argsort(cells, kind="stable")
    assert np.array_equal(S, X[order])
    engine.set_points(X[::-1].copy())                   # a new cloud invalidates the cached order
    assert np.array_equal(engine.sorted_points(), X[::-1][np.argsort(cells[::-1], kind="stable")])


def test_flat_sorted_sweep_against_the_unsorted_one(engine, bun000):
    """configs[1] through em_flat7_kernel (moments about each component's mean, file order) and em_flat8_kernel (about one
    origin per CTA, cell order): both within the tolerance of the float64 oracle, and of each other"""
    from oracle import c_oracle
    X, J = bun000, 800
    mu0 = X[np.random.default_rng(1).choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * np.float32(1e-4), (J, 1, 1))
    engine.set_points(X)
    w0 = np.full(J, 1.0 / J, np.float32)
    r7 = engine.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, tile_points=8)
    r8 = engine.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, tile_points=9)
    r9 = engine.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=10, tile_points=10)
    ow, omu, ocov, oll = c_oracle.flat_fit(X, mu0, 10, np.float32(1e-4))
    for name, r in (("em_flat7", r7), ("em_flat8", r8), ("em_flat8 cholesky-form densities", r9)):
        errs = (rel_fro(r["weights"], ow), rel_fro(r["means"], omu), rel_fro(r["covs"], ocov), rel_fro(r["ll"], oll))
        print("%s rel_fro (w, mu, cov, ll) = %.2e %.2e %.2e %.2e" % ((name,) + errs))
        assert max(errs) < TOL, (name, errs)
    assert rel_fro(r8["means"], r7["means"]) < TOL and rel_fro(r8["covs"], r7["covs"]) < TOL
    assert not np.array_equal(r8["covs"], r7["covs"])          # the switch really selects two kernels


@pytest.mark.parametrize("J", [1, 5, 31, 32, 33, 64, 100, 129, 160, 161, 512, 544, 545])
def test_flat_component_count_sweep(engine, bun000, J):
    """every slot/group configuration of the sweep kernel (ragged J, 1..16 warps per group)"""
    from oracle import flat_gmm
    X = bun000[::3]
    rng = np.random.default_rng(J)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 3e-4, (J, 1, 1))
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=3)
    ow, omu, ocov, oll = flat_gmm.cpp_fit(X, mu0, 3, sigma0_sq=np.float32(3e-4))
    assert rel_fro(r["means"], omu) < TOL and rel_fro(r["covs"], ocov) < TOL and rel_fro(r["weights"], ow) < TOL
    assert rel_fro(r["ll"], oll) < TOL


def test_flat_run_to_run_deterministic(engine, bun000):
    """partial rows + fixed-order fp64 reduction: two runs are bit-identical"""
    X = bun000
    J = 800
    rng = np.random.default_rng(1)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1))
    engine.set_points(X)
    a = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=10)
    b = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=10)
    assert (a["means"] == b["means"]).all() and (a["covs"] == b["covs"]).all() and (a["weights"] == b["weights"]).all()
    assert (a["ll"] == b["ll"]).all()


def test_flat_full_sigma_bug_flag(engine, bun000):
    from oracle import flat_gmm
    X = bun000[::20]
    J = 8
    rng = np.random.default_rng(3)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32), (J, 1, 1))
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=3, sigma_bug=True)
    ow, omu, ocov, _ = flat_gmm.cpp_fit(X, mu0, 3, sigma0_sq=1.0, sigma_bug=True)
    assert rel_fro(r["means"], omu) < TOL and rel_fro(r["covs"], ocov) < TOL


def test_flat_full_size_properties(engine, bun000):
    """full bun000 x J=800 (config 2): properties that need no oracle"""
    X = bun000
    J = 800
    rng = np.random.default_rng(1)
    mu0 = X[rng.choice(len(X), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1))
    engine.set_points(X)
    r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=10)
    assert abs(float(r["weights"].astype(np.float64).sum()) - 1.0) < 1e-5
    assert np.all(np.diff(r["ll"]) > -1e-3 * abs(r["ll"][0]))          # EM monotonicity
    # mixture mean == data mean (first-moment conservation of any EM step)
    mm = (r["weights"][:, None].astype(np.float64) * r["means"]).sum(0)
    assert np.allclose(mm, X.astype(np.float64).mean(0), atol=1e-6)
    ev = np.linalg.eigvalsh(r["covs"].astype(np.float64))
    assert (ev > 0).all()
    assert np.abs(r["covs"] - np.swapaxes(r["covs"], 1, 2)).max() == 0
    # determinism of everything but atomic order: a second run agrees to fp32 noise
    r2 = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=10)
    assert rel_fro(r2["means"], r["means"]) < 1e-6


def test_flat_edge_cases(engine):
    import hgmm_b200
    X = np.random.default_rng(0).normal(size=(5, 3)).astype(np.float32)
    engine.set_points(X)
    # ragged: fewer points than a tile, J not a multiple of 32, one component
    r = engine.fit_flat(X[:1], np.eye(3, dtype=np.float32)[None], np.ones(1, np.float32), cov_type="full", max_iter=2)
    assert np.allclose(r["means"][0], X.mean(0), atol=1e-6) and abs(r["weights"][0] - 1) < 1e-6
    assert np.allclose(r["covs"][0], np.cov(X.T, bias=True), atol=1e-5)
    with pytest.raises(hgmm_b200.HgmmError):
        engine.fit_flat(np.zeros((2000, 3)), np.tile(np.eye(3), (2000, 1, 1)), np.ones(2000) / 2000, cov_type="full", max_iter=1)
    with pytest.raises(hgmm_b200.HgmmError):
        hgmm_b200.Engine(0).fit_flat(X[:1], np.eye(3)[None], np.ones(1), cov_type="full", max_iter=1)     # no points set


# ------------------------------------------------------------------ tree build (C3 semantics at oracle sizes)
@pytest.mark.parametrize("tag", ["bun600_L2", "bun1500_L2"])
def test_tree_build_matches_reference_golden(engine, tag):
    g = gold("tree_build_%s.npz" % tag)
    engine.set_points(g["points"])
    r = engine.fit_tree(g["init_means"], int(g["L"]), ls=float(g["ls"]), ld=float(g["ld"]), sig2=float(g["sig2"]), ll_mode="level")
    assert list(r["iters"]) == list(g["oracle_iters"])
    assert rel_fro(r["pi"], g["ref_pi"]) < TOL
    assert rel_fro(r["mu"], g["ref_mu"]) < TOL
    assert rel_fro(r["cov"], g["ref_cov"]) < TOL
    assert (r["current"] == g["oracle_current"]).mean() > 0.995


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("ll_mode", ["level", "estep"])
@pytest.mark.parametrize("L,n", [(2, 5000), (3, 20000)])
def test_tree_build_matches_oracle(engine, L, n, ll_mode, variant):
    from oracle import hgmm_tree, synth
    X = synth.bunny_like(n, seed=7)
    init = X[hgmm_tree.reference_init_indices(L)]
    engine.set_points(X)
    r = engine.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode=ll_mode, variant=variant)
    opi, omu, ocov, ocur, oit, _ = hgmm_tree.build_gmm_tree(X, L, 20.0, 1e-4, init.astype(np.float64), sig2=np.float32(4e-4),
                                                          ll_mode=ll_mode, return_trace=True)
    assert list(r["iters"]) == list(oit)
    assert rel_fro(r["pi"], opi) < TOL
    assert rel_fro(r["mu"], omu) < TOL
    assert rel_fro(r["cov"], ocov) < 3 * TOL
    assert (r["current"] == ocur).mean() > 0.995


def test_tree_lidar_properties(engine):
    """config 3 shape (100k-point synthetic LiDAR sweep, L=4): oracle-free invariants"""
    from oracle import hgmm_tree, synth
    X = synth.lidar_sweep(100000, seed=2024)
    L = 4
    init = X[hgmm_tree.reference_init_indices(L)]
    engine.set_points(X)
    r = engine.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=25.0, ll_mode="estep")
    nt = hgmm_tree.n_total(L)
    assert r["pi"].shape == (nt,)
    prev = 1.0 + 1e-5
    for l in range(L):
        lb, le = hgmm_tree.level(l), hgmm_tree.level(l + 1)
        s = float(r["pi"][lb:le].astype(np.float64).sum())
        # the reference semantics drop a point whose 8 child densities all fall below 1e-15 (hgmm_gpu.py:404-406),
        # so level mass can only shrink with depth and never exceed 1
        assert s <= prev + 1e-5 and s > 0.5, (l, s)
        prev = s
    assert float(r["pi"][:8].astype(np.float64).sum()) > 0.99
    cur = r["current"]
    lb = hgmm_tree.level(L - 1)
    assert cur.min() >= lb and cur.max() < nt
    # hard-assignment histogram is consistent with the soft masses of the leaf level (same order of magnitude, same support)
    cnt = np.bincount(cur - lb, minlength=nt - lb)
    live = r["pi"][lb:] > 0
    assert (cnt[~live] == 0).mean() > 0.9      # dead points (all 8 densities < 1e-15) fall into slot 0, even a blank one
    # children stay inside their parent: leaf means are closer to their own parent's mean than to a random parent
    par = (np.arange(lb, nt) // 8) - 1
    d_own = np.linalg.norm(r["mu"][lb:][live] - r["mu"][par][live], axis=1)
    d_rand = np.linalg.norm(r["mu"][lb:][live] - r["mu"][np.roll(par, 777)][live], axis=1)
    assert np.median(d_own) < 0.5 * np.median(d_rand)
    ev = np.linalg.eigvalsh(r["cov"][lb:][live].astype(np.float64))
    assert (ev[:, 2] > 0).all()


def test_buildGMMTree_api_is_drop_in(engine):
    from hgmm_b200 import hgmm
    g = gold("tree_build_bun600_L2.npz")
    pi, mu, cov = hgmm.buildGMMTree(g["points"], 2, 20.0, 1e-4, sig2=float(g["sig2"]), init_means=g["init_means"], engine=engine)
    assert rel_fro(pi, g["ref_pi"]) < TOL and rel_fro(mu, g["ref_mu"]) < TOL and rel_fro(cov, g["ref_cov"]) < TOL


# ------------------------------------------------------------------ registration (C4 semantics)
def test_registration_steps_match_reference_golden(engine):
    g = gold("tree_reg_bun1500_L2.npz")
    L = int(g["L"])
    engine.tree_set_model(L, g["pi"], g["mu"], g["cov"])
    engine.reg_set_target(g["target"])
    m0, m1, m2 = engine.reg_estep(np.identity(3), np.zeros(3), float(g["lambda_c"]), len(g["pi"]))
    assert rel_fro(m0, g["ref_M0"]) < TOL
    assert rel_fro(m1, g["ref_M1"]) < TOL
    rot, t, q = engine.reg_mstep(np.identity(3), np.zeros(3), "twist_lstsq")
    assert rel_fro(rot, g["ref_step_rot"]) < TOL
    assert rel_fro(t, g["ref_step_t"]) < 10 * TOL
    assert abs(q - float(g["ref_step_q"][0])) < 1e-3 * abs(float(g["ref_step_q"][0]))


def test_registration_loop_matches_reference_golden(engine):
    g = gold("tree_reg_bun1500_L2.npz")
    L = int(g["L"])
    engine.tree_set_model(L, g["pi"], g["mu"], g["cov"])
    engine.reg_set_target(g["target"])
    rot, t, q, it, hist = engine.register_tree(solver="twist_lstsq", maxiter=20, tol=1e-4, lambda_c=float(g["lambda_c"]))
    inv = np.c_[rot.T, -rot.T @ t]
    ref = np.c_[g["ref_rot"], g["ref_t"]]
    assert rel_fro(inv, ref) < TOL
    assert abs(it - int(g["oracle_iters"])) <= 1


def test_registration_gmmtree_api(engine):
    from hgmm_b200 import hgmm
    g = gold("tree_reg_bun1500_L2.npz")
    gt = hgmm.GMMTree(None, tree_level=int(g["L"]), lambda_c=float(g["lambda_c"]), engine=engine)
    gt.set_model(g["pi"], g["mu"], g["cov"])
    res = gt.registration(g["target"], 20, 1e-4)
    assert rel_fro(res.transformation.rot, g["ref_rot"]) < TOL and rel_fro(res.transformation.t, g["ref_t"]) < 10 * TOL
    est = gt.expectation_step(g["target"])
    assert rel_fro(est.momentZero, g["ref_M0"]) < TOL
    ms = gt.maximization_step(est, hgmm.RigidTransformation())
    assert rel_fro(ms.transformation.rot, g["ref_step_rot"]) < TOL


def test_registration_procrustes_matches_oracle(engine):
    from oracle import registration as oreg
    g = gold("tree_reg_bun1500_L2.npz")
    L = int(g["L"])
    engine.tree_set_model(L, g["pi"], g["mu"], g["cov"])
    engine.reg_set_target(g["target"])
    m0, m1, _ = engine.reg_estep(np.identity(3), np.zeros(3), float(g["lambda_c"]), len(g["pi"]), want_m2=False)
    rot, t, q = engine.reg_mstep(np.identity(3), np.zeros(3), "procrustes_svd")
    oR, ot, oq = oreg.reg_m_step_procrustes(m0, m1, g["mu"].astype(np.float32).astype(np.float64), np.identity(3), np.zeros(3))
    assert rel_fro(rot, oR) < TOL and np.abs(t - ot).max() < 1e-6
    assert abs(q - oq) < 1e-6 * max(abs(oq), 1e-12) + 1e-12
    assert abs(np.linalg.det(rot) - 1.0) < 1e-9 and np.abs(rot @ rot.T - np.identity(3)).max() < 1e-9
    # the whole loop against the oracle's loop (same solver), ONE direction: the engine returns the forward transform
    # (target -> model), the oracle / reference its inverse (hgmm_gpu.py:768)
    rot, t, q, it, _ = engine.register_tree(solver="procrustes_svd", maxiter=30, tol=1e-9)
    lR, lt, lq, lit = oreg.registration(g["target"], g["pi"].astype(np.float64), g["mu"].astype(np.float32).astype(np.float64),
                                        g["cov"].astype(np.float32).astype(np.float64), L, float(g["lambda_c"]), 30, 1e-9,
                                        solver="procrustes")
    assert rel_fro(np.c_[rot.T, -rot.T @ t], np.c_[lR, lt]) < 10 * TOL, (rel_fro(rot.T, lR), it, lit)
    assert abs(it - lit) <= 1
    # target = true_rot . source + t0, so the forward rotation is true_rot^T (8 degrees about z): direction check
    assert rel_fro(rot, g["true_rot"].T) < 5e-2 and rel_fro(rot, g["true_rot"]) > 0.1      # oracle: 3.9e-2 / 0.19


def test_bunny_registration_matches_oracle_on_real_scans(engine, bun000, bun045):
    """config 4 inputs (bun000 -> bun045, data/bun.conf pose 34.3 deg about y): the engine's registration loop against
    the oracle's on the real partial-overlap scans.  The reference algorithm itself does not recover the bun.conf pose
    here (the oracle drifts to ~8 deg even when started AT the true pose), so ground truth is reported, not asserted."""
    from oracle import hgmm_tree, registration as oreg
    q = np.array([0.00548449, -0.294635, -0.0038555, 0.955586])     # x y z w, data/bun.conf:3
    x, y, z, w = q
    Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                   [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                   [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    S, T = bun000[::6].astype(np.float64), bun045[::6].astype(np.float64)
    L = 2
    init = S[hgmm_tree.reference_init_indices(L)]
    pi, mu, cov, _ = hgmm_tree.build_gmm_tree(S, L, 20.0, 1e-4, init, sig2=4e-4, ll_mode="estep")
    pi, mu, cov = pi.astype(np.float32), mu.astype(np.float32), cov.astype(np.float32)
    th = np.deg2rad(25.0)
    R0 = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    # p_bun000frame = Rq^T p_bun045 + tq: the forward transform of the target is ~ (Rq^T, tq)
    oRinv, otinv, oq, oit = oreg.registration(T, pi.astype(np.float64), mu.astype(np.float64), cov.astype(np.float64), L, 0.01, 12, 1e-9,
                                             rot=R0, t=np.zeros(3))
    engine.tree_set_model(L, pi, mu, cov)
    engine.reg_set_target(T)
    rot, t, qq, it, hist = engine.register_tree(rot=R0, t=np.zeros(3), solver="twist_lstsq", maxiter=12, tol=1e-9, lambda_c=0.01)
    assert it == oit == 12
    assert rel_fro(np.c_[rot.T, -rot.T @ t], np.c_[oRinv, otinv]) < 10 * TOL       # 12 chained fp32 E-steps on real scans
    ang = np.rad2deg(np.arccos(np.clip((np.trace(rot @ Rq) - 1) / 2, -1, 1)))
    print("angle to the bun.conf pose after 12 iterations: %.2f deg (oracle: same algorithm)" % ang)


@pytest.mark.parametrize("solver", ["procrustes_svd", "twist_lstsq"])
def test_flat_registration_matches_oracle(engine, bun000, solver):
    """configs[3] semantics (flat J=100 mixture + weighted-Procrustes / twist solve): the device loop against the float64
    restatement on a synthetic rigid motion of the real scan; the forward transform must also be the true one."""
    from oracle import registration as oreg
    S = bun000[::4]
    J = 100
    rng = np.random.default_rng(11)
    mu0 = S[rng.choice(len(S), J, replace=False)]
    engine.set_points(S)
    fit = engine.fit_flat(mu0, np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1)), np.full(J, 1 / J, np.float32),
                          cov_type="full", max_iter=10)
    th = np.deg2rad(8.0)
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    T = (bun000[1::4].astype(np.float64) @ Rz.T + np.array([0.004, -0.003, 0.002])).astype(np.float32)
    engine.reg_set_target(T)
    # stopping thresholds well above the fp32 noise of q (Procrustes q ~ 3e-3, twist q ~ 125 here)
    tol = 1e-6 if solver == "procrustes_svd" else 1e-2
    rot, t, q, it, hist = engine.register_flat(solver=solver, maxiter=30, tol=tol)
    oR, ot, oq, oit = oreg.flat_registration(T, fit["weights"], fit["means"], fit["covs"], 30, tol,
                                             solver="twist_lstsq" if solver == "twist_lstsq" else "procrustes")
    assert abs(it - oit) <= 1, (it, oit)
    assert rel_fro(np.c_[rot.T, -rot.T @ t], np.c_[oR, ot]) < 10 * TOL        # <= 30 chained fp32 E-steps
    assert abs(q - oq) < 1e-3 * abs(oq) + 1e-9
    # the target is Rz . source + t0: the forward transform (target -> model) is Rz^T, one direction only
    assert rel_fro(rot, Rz.T) < 2e-2 and rel_fro(rot, Rz) > 0.1
    assert abs(np.linalg.det(rot) - 1.0) < 1e-9


def test_flat_registration_config4_real_scans(engine, bun000, bun045):
    """configs[3] as stated: flat J=100 fit of bun000 (10 EM iterations) + weighted-Procrustes registration of bun045, the
    engine's loop against the oracle's on the real partial-overlap scans; the angle to the data/bun.conf pose is reported."""
    from oracle import registration as oreg
    J = 100
    rng = np.random.default_rng(12)
    mu0 = bun000[rng.choice(len(bun000), J, replace=False)]
    engine.set_points(bun000)
    fit = engine.fit_flat(mu0, np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1)), np.full(J, 1 / J, np.float32),
                          cov_type="full", max_iter=10)
    T = bun045[::3]
    engine.reg_set_target(T)
    rot, t, q, it, hist = engine.register_flat(solver="procrustes_svd", maxiter=20, tol=1e-4)
    oR, ot, oq, oit = oreg.flat_registration(T, fit["weights"], fit["means"], fit["covs"], 20, 1e-4, solver="procrustes")
    assert abs(it - oit) <= 1, (it, oit)
    assert rel_fro(np.c_[rot.T, -rot.T @ t], np.c_[oR, ot]) < 10 * TOL
    x, y, z, w = 0.00548449, -0.294635, -0.0038555, 0.955586          # data/bun.conf:3
    Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                   [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                   [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    ang = np.rad2deg(np.arccos(np.clip((np.trace(rot @ Rq) - 1) / 2, -1, 1)))
    print("flat J=100 Procrustes, %d iterations: angle to the bun.conf pose %.2f deg" % (it, ang))
    # run to convergence with the twist solver the same registration RECOVERS the scanner's pose (data/bun.conf:3: 34.3 degrees
    # about y, p_bun000 = Rq^T p_bun045 + tq): the accuracy anchor of configs[3] (the float64 oracle reaches 0.31 degrees)
    rot, t, q, it, hist = engine.register_flat(solver="twist_lstsq", maxiter=150, tol=1e-2)
    ang = np.rad2deg(np.arccos(np.clip((np.trace(rot @ Rq) - 1) / 2, -1, 1)))
    tq = np.array([-0.0520211, -0.000383981, -0.0109223])
    print("flat J=100 twist, %d iterations: angle to the bun.conf pose %.2f deg, |t - tq| = %.4f m" % (it, ang, np.linalg.norm(t - tq)))
    assert ang < 1.0 and np.linalg.norm(t - tq) < 3e-3


def _lidar_cloud(tag):
    from hgmm_b200 import synth
    if tag == "lidar100k_L4":
        return synth.lidar_sweep(100000, seed=2024)
    P = synth.lidar_sweep(1000000, seed=2025)          # the first 50k points of the seeded shuffle (oracle/make_golden_lidar.py)
    return P[np.random.default_rng(0).permutation(len(P))[:50000]]


def test_adaptive_tree_build_matches_oracle(engine, bun000):
    """the pruned (ragged) build -- nodes that are blank, flat enough (complexity <= lambda_c) or too light
    (< min_points) become terminal after their level, their subtree stays blank -- against the oracle's restatement of the same
    rule; the tree must be genuinely ragged, and registration must run on it."""
    from oracle import hgmm_tree
    from hgmm_b200 import hgmm as H
    X = bun000[::2]
    L = 3
    init = X[hgmm_tree.reference_init_indices(L)]
    kw = dict(prune_lambda_c=0.02, prune_min_points=40.0)
    engine.set_points(X)
    r = engine.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep", **kw)
    opi, omu, ocov, ocur, oit, _ = hgmm_tree.build_gmm_tree(X, L, 20.0, 1e-4, init.astype(np.float64), sig2=np.float32(4e-4), ll_mode="estep",
                                                            return_trace=True, **kw)
    assert r["iters"].tolist() == list(oit)
    assert rel_fro(r["pi"], opi) < TOL and rel_fro(r["mu"], omu) < TOL and rel_fro(r["cov"], ocov) < 3 * TOL
    assert float((r["current"] == ocur).mean()) > 0.9995
    full = engine.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep")
    lb = hgmm_tree.level(L - 1)
    live_pruned, live_full = int((r["pi"][lb:] > 0).sum()), int((full["pi"][lb:] > 0).sum())
    print("live leaves: pruned %d, full %d of %d" % (live_pruned, live_full, hgmm_tree.n_total(L) - lb))
    assert live_pruned < 0.8 * live_full                                    # the tree is ragged ...
    deep = H.deepest_live_node(r["current"], r["pi"], L)
    assert (r["pi"][deep] > 0).all() and (deep < lb).mean() > 0.1           # ... a good share of the points ends above the leaf level
    # the ragged frontier (deepest live node of every point) carries the cloud's mass once -- up to the soft responsibilities
    # that leak to siblings and the points too far from every child to be counted at all (den <= 1e-15)
    front = np.zeros(len(r["pi"]), bool)
    front[np.unique(deep)] = True
    assert 0.9 < float(r["pi"][front].sum()) < 1.05
    with pytest.raises(Exception):
        engine.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="level", **kw)     # needs the persistent (estep) path
    th = np.deg2rad(5.0)
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    engine.tree_set_model(L, r["pi"], r["mu"], r["cov"])
    engine.reg_set_target((X[::2] @ Rz.T + np.array([0.002, 0.001, -0.001])).astype(np.float32))
    rot, t, q, it, _ = engine.register_tree(solver="twist_lstsq", maxiter=20, tol=1e-6, lambda_c=0.02)
    assert rel_fro(rot, Rz.T) < 2e-2


@pytest.mark.parametrize("force", [1, 2, 4, 7])
def test_tree_level_kernel_overflow_paths(engine, bun000, force):
    """the persistent level kernel keeps its points, chunk descriptors and fold results on chip -- up to capacities that clouds of
    ordinary size never exceed.  HGMM_TREE_FORCE shrinks them (bit 0: points streamed from L2, bit 1: descriptors from global
    memory, bit 2: a two-slot fold stage, the rest goes straight to L2 atomics): the fallback paths must give the same tree."""
    from oracle import hgmm_tree
    X = bun000[::2]
    L = 3
    init = X[hgmm_tree.reference_init_indices(L)]
    engine.set_points(X)
    kw = dict(ls=0.0, ld=1e-4, sig2=4e-4, ll_mode="estep", max_iters_per_level=10)
    ref = engine.fit_tree(init, L, **kw)
    os.environ["HGMM_TREE_FORCE"] = str(force)
    try:
        r = engine.fit_tree(init, L, **kw)
    finally:
        del os.environ["HGMM_TREE_FORCE"]
    assert r["iters"].tolist() == ref["iters"].tolist()
    # same arithmetic per point, another summation order (fp32 partial sums, fp64 atomics): rounding-level agreement at the root,
    # 1e-4 on the weights everywhere, and on the nodes that carry mass (a massless node may flip blank / alive on the last bit)
    assert rel_fro(r["mu"][:8], ref["mu"][:8]) < 1e-5 and rel_fro(r["cov"][:8], ref["cov"][:8]) < 1e-5
    for lv in range(L):
        e = _tree_level_errors(r, ref, lv)
        assert e["mu_w"] < 3e-3 and e["cov_w"] < 3e-3 and e["dpi_l1"] < 3e-3, (lv, e)
    assert float((r["current"] == ref["current"]).mean()) > 0.998


def test_tree_more_nodes_than_points(engine, bun000):
    """ragged extreme: 300 points under a depth-3 tree (584 nodes, 512 leaves): most nodes stay blank, nothing may go NaN, and the
    result must still equal the oracle's"""
    from oracle import hgmm_tree
    X = bun000[::134][:300]
    L = 3
    init = np.resize(X, (hgmm_tree.n_total(L), 3))[hgmm_tree.reference_init_indices(L) % len(X)]
    engine.set_points(X)
    r = engine.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep")
    opi, omu, ocov, ocur, oit, _ = hgmm_tree.build_gmm_tree(X, L, 20.0, 1e-4, init.astype(np.float64), sig2=np.float32(4e-4), ll_mode="estep",
                                                            return_trace=True)
    assert np.isfinite(r["mu"]).all() and np.isfinite(r["cov"]).all()
    assert r["iters"].tolist() == list(oit)
    assert rel_fro(r["pi"], opi) < TOL and float((r["current"] == ocur).mean()) == 1.0
    live = (opi > 0) & (r["pi"] > 0)
    assert rel_fro(r["mu"][live], omu[live]) < TOL and rel_fro(r["cov"][live], ocov[live]) < 1e-3      # single-point nodes: Sigma = 0 +- rounding


def _tree_level_errors(r, g, lv):
    """node-wise distances of level `lv` between a fitted tree r and the oracle fixture g, on nodes alive in both:
    |d mu| / sqrt(tr Sigma) and |d Sigma|_F / |Sigma|_F per node -> (mass-weighted means, medians, |d pi|_1, mass of the nodes
    alive in only one of the two).  Why not one unweighted Frobenius norm over all nodes: a node whose zeroth moment sits at the
    reference's blanking threshold (M0 < ld = 1e-4 of a POINT) is blank (mu = 0) or alive (|mu| ~ 50 m) by the last bit of M0,
    and a single such massless node then dominates ||d mu||_F (observed: one of 64 -> 0.28) while carrying 1e-9 of the mass."""
    from oracle import hgmm_tree
    a, b = hgmm_tree.level(lv), hgmm_tree.level(lv + 1)
    pg, pr = g["pi"][a:b].astype(np.float64), r["pi"][a:b].astype(np.float64)
    both = (pg > 0) & (pr > 0)
    w = pg[both]
    dm = np.linalg.norm(r["mu"][a:b].astype(np.float64) - g["mu"][a:b], axis=1) / np.sqrt(np.maximum(np.trace(g["cov"][a:b], axis1=1, axis2=2), 1e-30))
    gc = g["cov"][a:b].reshape(-1, 9).astype(np.float64)
    dc = np.linalg.norm(r["cov"][a:b].reshape(-1, 9).astype(np.float64) - gc, axis=1) / np.maximum(np.linalg.norm(gc, axis=1), 1e-30)
    nc = np.linalg.norm(gc, axis=1)
    ok = both & (nc > 1e-9) & (nc < 1e6)                      # (single-point nodes have Sigma = 0: nothing to normalise by)
    return {"mu_w": float((pg[ok] * dm[ok]).sum() / pg[ok].sum()), "cov_w": float((pg[ok] * dc[ok]).sum() / pg[ok].sum()),
            "mu_med": float(np.median(dm[ok])), "cov_med": float(np.median(dc[ok])), "dpi_l1": float(np.abs(pr - pg).sum()),
            "one_sided_mass": float(pg[~both].sum() + pr[~both].sum()), "w": float(w.sum())}


@pytest.mark.parametrize("tag,L", [("lidar100k_L4", 4), ("lidar50k_L5", 5)])
def test_tree_config_size_matches_oracle_golden(engine, tag, L):
    """configs[2] at its full size (100k-point sweep, depth 4: 4680 nodes) and the 50k-point / depth-5 subsample of configs[4]
    (37448 nodes, 32768 leaves, ~1.5 points per leaf: the near-empty-node regime) against the float64 oracle's fixture: every
    level's E-step, M-step and partition at config size, two EM iterations per level.  Held to: 1e-6 at the root level and 1e-4
    at level 1 (every node that carries mass; observed 2e-7 ... 1e-5 from run to run), and below the root a mass-weighted node error <= 3e-3 with a MEDIAN node error
    <= 1e-5, >= 99.8 % of the points in the oracle's leaf.  The residual below level 1 is not arithmetic noise in the moments: it
    is the reference's own blanking rule (M0 < ld) and dead-point rule (all eight densities below 1e-15 -> child 0) flipping on
    the last bits for massless nodes / far points (see _tree_level_errors), which moves a few dozen of the 100 000 points to
    another leaf -- both of this library's tree kernels, and the reference's float32 arrays, sit at the same distance."""
    g = gold("tree_build_%s_estep_fixed2.npz" % tag)
    P = _lidar_cloud(tag)
    assert np.allclose(np.asarray(P, np.float64).sum(axis=0), g["cloud_checksum"], rtol=0, atol=1e-6 * len(P)), "the generator drifted"
    from oracle import hgmm_tree
    init = P[hgmm_tree.reference_init_indices(L)]
    engine.set_points(P)
    r = engine.fit_tree(init, L, ls=0.0, ld=float(g["ld"]), sig2=float(g["sig2"]), ll_mode="estep", max_iters_per_level=2)
    assert r["iters"].tolist() == g["iters"].tolist() == [2] * L
    a, b = hgmm_tree.level(0), hgmm_tree.level(1)
    assert rel_fro(r["pi"][a:b], g["pi"][a:b]) < 1e-6 and rel_fro(r["mu"][a:b], g["mu"][a:b]) < 1e-6 and rel_fro(r["cov"][a:b], g["cov"][a:b]) < 1e-6
    for lv in range(L):
        e = _tree_level_errors(r, g, lv)
        print("level %d: %s" % (lv, {k: "%.1e" % v for k, v in e.items()}))
        lim = 1e-6 if lv == 0 else 1e-4 if lv == 1 else 3e-3
        assert e["mu_w"] < lim and e["cov_w"] < lim, (lv, e)
        assert e["mu_med"] < 2e-5 and e["cov_med"] < 2e-5, (lv, e)
        assert e["dpi_l1"] < 3e-3 and e["one_sided_mass"] < 3e-3, (lv, e)
    lb = hgmm_tree.level(L - 1)
    agree = float(((r["current"] - lb) == g["current_leaf"].astype(np.int64)).mean())
    assert agree > 0.998, agree
    assert abs(r["q"][0] - g["q_last"][0]) < 1e-6 * abs(g["q_last"][0])          # root level: q itself to 1e-6
    assert abs(r["q"][-1] - g["q_last"][-1]) < 2e-3 * abs(g["q_last"][-1])       # leaf level: a flipped dead point costs log(1e-15)


@pytest.mark.parametrize("tag,L,fixed", [("lidar100k_L4", 4, 12), ("lidar50k_L5", 5, 10)])
def test_tree_config_size_long_run_stays_within_the_fp32_storage_envelope(engine, tag, L, fixed):
    """the same builds at 12 / 10 iterations per level: the tree build amplifies ANY perturbation through its hard hand-offs
    (test_oracle_golden.py::test_tree_fp32_storage_alone_moves_the_config_size_build: float32 parameter storage alone moves the
    float64 oracle by 6e-4 / 4e-3 / 7e-4 here), so the long run is held to that envelope: 1e-5 at the root level, mass-weighted
    node errors <= 5e-2 / 1e-1 below it, >= 98 % of the points in the oracle's leaf, the leaf level's log-likelihood to 1e-2."""
    g = gold("tree_build_%s_estep_fixed%d.npz" % (tag, fixed))
    P = _lidar_cloud(tag)
    from oracle import hgmm_tree
    init = P[hgmm_tree.reference_init_indices(L)]
    engine.set_points(P)
    r = engine.fit_tree(init, L, ls=0.0, ld=float(g["ld"]), sig2=float(g["sig2"]), ll_mode="estep", max_iters_per_level=fixed)
    assert r["iters"].tolist() == [fixed] * L
    assert rel_fro(r["mu"][:8], g["mu"][:8]) < 1e-5 and rel_fro(r["cov"][:8], g["cov"][:8]) < 1e-5 and rel_fro(r["pi"][:8], g["pi"][:8]) < 1e-5
    for lv in range(1, L):
        e = _tree_level_errors(r, g, lv)
        print("level %d: %s" % (lv, {k: "%.1e" % v for k, v in e.items()}))
        assert e["mu_w"] < 5e-2 and e["cov_w"] < 1e-1 and e["dpi_l1"] < 3e-2, (lv, e)
    agree = float(((r["current"] - hgmm_tree.level(L - 1)) == g["current_leaf"].astype(np.int64)).mean())
    assert agree > 0.98, agree
    assert abs(r["q"][-1] - g["q_last"][-1]) < 1e-2 * abs(g["q_last"][-1])        # one flipped dead point costs log(1e-15) = 34.5


@pytest.mark.parametrize("tag,L", [("lidar100k_L4", 4), ("lidar50k_L5", 5)])
def test_tree_config_size_converged_against_oracle_golden(engine, tag, L):
    """the same workloads run to the reference's stopping rule (|q - prevQ| < 20).  The rule sits on a plateau of q: WHICH
    iteration crosses it depends on the last bits of q (the reference's own fp32 atomics make its count non-deterministic run to
    run), so the counts are held to the oracle's at the two top levels only; the converged log-likelihood must agree to 2 %."""
    g = gold("tree_build_%s_estep.npz" % tag)
    P = _lidar_cloud(tag)
    from oracle import hgmm_tree
    init = P[hgmm_tree.reference_init_indices(L)]
    engine.set_points(P)
    r = engine.fit_tree(init, L, ls=float(g["ls"]), ld=float(g["ld"]), sig2=float(g["sig2"]), ll_mode="estep")
    print("iterations", r["iters"].tolist(), "oracle", g["iters"].tolist())
    assert np.abs(r["iters"][:2].astype(np.int64) - g["iters"][:2].astype(np.int64)).max() <= 1
    assert rel_fro(r["mu"][:8], g["mu"][:8]) < TOL and rel_fro(r["cov"][:8], g["cov"][:8]) < TOL
    assert abs(r["q"][-1] - g["q_last"][-1]) < 2e-2 * abs(g["q_last"][-1])
    assert abs(float(r["pi"][hgmm_tree.level(L - 1):].sum()) - float(g["pi"][hgmm_tree.level(L - 1):].sum())) < 2e-2


def test_flat_far_points_take_the_exact_path(engine):
    """points > 11 sigma from every component underflow the fixed-reference sums; the sweep must fall back to the exact
    per-point maximum and still match the (max-shifted) oracle.  (Outliers are kept within ~60 sigma: beyond that |log2 p|
    itself exceeds 1e5 and fp32 cannot resolve the split between neighbouring components to 1e-4 -- nor can the reference.)"""
    from oracle import flat_gmm
    rng = np.random.default_rng(4)
    X = np.concatenate([rng.normal(0, 0.01, (500, 3)), rng.normal(0, 0.01, (500, 3)) + [1.0, 0, 0],
                        [[0.5, 0.3, 0.0], [0.45, -0.2, 0.1], [0.3, 0.3, 0.3]]]).astype(np.float32)      # 30-60 sigma outliers
    for J in (2, 40):
        mu0 = np.concatenate([X[:J // 2], X[500:500 + J - J // 2]])
        cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1))
        engine.set_points(X)
        r = engine.fit_flat(mu0, cov0, np.full(J, 1 / J, np.float32), cov_type="full", max_iter=2)
        ow, omu, ocov, oll = flat_gmm.cpp_fit(X, mu0, 2, sigma0_sq=np.float32(1e-4))
        assert rel_fro(r["weights"], ow) < TOL and rel_fro(r["means"], omu) < TOL
        assert rel_fro(r["covs"], ocov) < TOL and rel_fro(r["ll"], oll) < TOL
        assert abs(float(r["weights"].sum()) - 1.0) < 1e-5


def test_fill_vbo(engine):
    import torch
    X = np.arange(30, dtype=np.float32).reshape(10, 3)
    engine.set_points(X)
    engine.reg_set_target(X[:4] + 100)
    pos = torch.zeros(14 * 4, device="cuda")
    col = torch.zeros(14 * 4, device="cuda")
    engine.fill_vbo(pos.data_ptr(), col.data_ptr(), 0.1)
    p = pos.cpu().numpy().reshape(14, 4)
    c = col.cpu().numpy().reshape(14, 4)
    assert np.allclose(p[:10, :3], -X / 0.1) and np.allclose(p[10:, :3], -(X[:4] + 100) / 0.1) and (p[:, 3] == 1).all()
    assert np.allclose(c[:10], [1.3, 1.3, 1.3, 1.0]) and np.allclose(c[10:], [1.3, 1.3, 0.3, 1.0])


def test_device_resident_input(engine, bun000):
    import torch
    Xd = torch.from_numpy(bun000[::10]).cuda()
    g = gold("flat_py_diag_sub4k_J8.npz")
    engine.set_points(Xd)
    r = engine.fit_flat(g["means0"], g["covs0"], g["weights0"], cov_type="diag", max_iter=10)
    assert rel_fro(r["means"], g["ref_means"]) < TOL


def test_reference_cuda_binary_pins_the_cpp_variant(engine, bun000, tmp_path):
    """the reference's OWN CUDA fitter (built from its sources by oracle/build_ref.sh) against the engine run with the
    same rand() draw and sigma_bug=1 (gmm_kernels.cu:97-103 reproduced): this is parity with the reference itself."""
    import ctypes
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gmm_cuda")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gmm_cuda not built")
    J = 8
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "bun000_xyz.npy"), str(J), "10", str(out)], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    raw = np.fromfile(out, dtype=np.float32)
    ref_mu, ref_w = raw[:3 * J].reshape(J, 3), raw[3 * J:]
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)                                   # process default; the reference never seeds (gmm_kernels.cu:375)
    idx = np.array([libc.rand() % len(bun000) for _ in range(J)])
    engine.set_points(bun000)
    g = engine.fit_flat(bun000[idx], np.tile(np.eye(3, dtype=np.float32), (J, 1, 1)), np.full(J, 1 / J, np.float32), cov_type="full",
                        max_iter=10, sigma_bug=True)
    assert rel_fro(g["means"], ref_mu) < TOL
    assert rel_fro(g["weights"], ref_w) < 1e-3      # the reference sums 40k fp32 terms serially per component


# ------------------------------------------------------------------ R5: L2-distance flat registration
@pytest.mark.parametrize("tag,which", [("sub4k_J50", "bun000"), ("b45sub4k_J50", "bun045")])
def test_flat_py_old_matches_reference_golden(engine, bun000, bun045, tag, which):
    from hgmm_b200 import gmmreg
    g = gold("flat_pyold_%s.npz" % tag)
    X = (bun000 if which == "bun000" else bun045)[::int(g["stride"])]
    inv, mu, w, cov, ll = gmmreg.train_gmm(X, 10, 0.0, g["means0"], g["covs0"], g["weights0"], engine=engine)
    assert rel_fro(mu, g["ref_means"]) < TOL and rel_fro(w, g["ref_weights"]) < TOL
    assert rel_fro(cov, g["ref_covs"]) < TOL and rel_fro(inv, g["ref_inv_cov"]) < TOL
    assert rel_fro(ll, g["ref_ll"]) < TOL and len(ll) == 10


def test_l2_cost_gradient_matches_reference_golden(engine):
    """float64 kernel vs the unmodified reference's RigidCostFunction at fixed theta (compare cost/gradient
    values, not optimiser iterates)."""
    g = gold("l2_cost.npz")
    engine.l2_set_mixtures(g["mu_s"], g["phi_s"], g["mu_t"], g["phi_t"])
    for th, rf, rg in zip(g["thetas"], g["ref_f"], g["ref_grad"]):
        f, grad = engine.l2_cost_grad(th, float(g["sigma"]))
        assert abs(f - rf) < 1e-11 * abs(rf)
        assert rel_fro(grad, rg) < 1e-9


def test_l2_cost_ragged_sizes_match_oracle(engine):
    from oracle import l2reg
    rng = np.random.default_rng(3)
    for Js, Jt in ((1, 1), (3, 700), (257, 5), (300, 300)):
        mu_s, mu_t = rng.normal(size=(Js, 3)) * 0.05, rng.normal(size=(Jt, 3)) * 0.05
        ps, pt = rng.uniform(0.1, 2.0, Js), rng.uniform(0.1, 2.0, Jt)
        th = np.r_[rng.normal(size=4), rng.normal(size=3) * 0.02]
        engine.l2_set_mixtures(mu_s, ps, mu_t, pt)
        f, grad = engine.l2_cost_grad(th, 0.04)
        of, og = l2reg.rigid_cost(th, mu_s, ps, mu_t, pt, 0.04)
        assert abs(f - of) < 1e-11 * abs(of) and rel_fro(grad, og) < 1e-9


def test_l2_scipy_loop_matches_reference_golden(engine):
    """the reference's own loop (SciPy BFGS, gmmreg.py:101-107) driving the device cost function, on the golden mixtures"""
    from hgmm_b200 import gmmreg
    c = gold("l2_cost.npz")
    g = gold("l2_reg_bunny_default.npz")

    class Fixed:
        calls = 0

        def init(self):
            pass

        def annealing(self):
            pass

        def compute(self, data):
            self.calls += 1
            return (c["mu_t"], c["phi_t"] / 1e3) if self.calls == 1 else (c["mu_s"], c["phi_s"] / 1e3)

    reg = gmmreg.L2DistRegistration(None, Fixed(), gmmreg.RigidCostFunction(engine=engine), sigma=float(g["sigma"]),
                                    use_estimated_sigma=False)
    tf = reg.registration(np.zeros((1, 3)), maxiter=1, tol=1e-3, opt_maxiter=10, opt_tol=1e-5)
    assert rel_fro(tf.rot, g["ref_rot"]) < 1e-3 and rel_fro(tf.t, g["ref_t"]) < 1e-2
    assert abs(reg.last_result["fun"] - float(g["oracle_f"])) < 1e-4 * abs(float(g["oracle_f"]))


def test_l2_device_bfgs_reaches_the_reference_minimum(engine):
    """one-launch BFGS.  Its iterates are not SciPy's (parity unpinned for the trajectory) and the reference's
    gradient is not the derivative of its cost (tests/test_oracle_golden.py), so two quasi-Newton runs stall at
    slightly different points of the same flat valley: the bar is the cost SciPy reaches on the oracle at the same sigma
    (within 1e-3) and a pose within a few degrees of it."""
    from oracle import l2reg
    from scipy.optimize import minimize
    c = gold("l2_cost.npz")
    sigma = float(gold("l2_reg_bunny_converged.npz")["sigma"])
    args = (c["mu_s"], c["phi_s"], c["mu_t"], c["phi_t"], sigma)
    engine.l2_set_mixtures(*args[:4])
    x0 = np.array([1.0, 0, 0, 0, 0, 0, 0])
    x, f, nit, nfev, status = engine.l2_optimize(x0, sigma, max_iter=200, gtol=1e-9)
    of, _ = l2reg.rigid_cost(x, *args)
    assert abs(f - of) < 1e-10 * abs(of)                      # the reported cost is the cost at the returned theta
    res = minimize(l2reg.rigid_cost, x0, args=args, method="BFGS", jac=True, tol=1e-9, options={"maxiter": 200})
    assert f <= res.fun + 1e-3 * abs(res.fun)
    dR = l2reg.quaternion_matrix3(x[:4]) @ l2reg.quaternion_matrix3(res.x[:4]).T
    assert np.rad2deg(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))) < 5.0
    assert np.abs(x[4:] - res.x[4:]).max() < 1e-2
    assert nit >= 1 and nfev >= nit and status in (0, 1, 2)
    # determinism: same launch, same answer
    x2, f2, *_ = engine.l2_optimize(x0, sigma, max_iter=200, gtol=1e-9)
    assert f2 == f and (x2 == x).all()


def test_registration_gmmreg_api_recovers_bunny_pose(engine, bun000, bun045):
    """registration_gmmreg(source, target) end to end (KMeans init on the host as the reference, both fits + cost on the
    device): bun000 -> bun045 is a 34 deg turn about y (data/bun.conf); 10 BFGS steps at the estimated sigma get within
    a few degrees, as the reference's own default run does (tests/golden/l2_reg_bunny_default.npz)."""
    from hgmm_b200 import gmmreg
    for opt in ("scipy", "device"):
        tf = gmmreg.registration_gmmreg(bun000[::10].astype(np.float64), bun045[::10].astype(np.float64), engine=engine,
                                        optimizer=opt)
        ang = np.rad2deg(np.arccos(np.clip((np.trace(tf.rot) - 1) / 2, -1, 1)))
        assert abs(np.linalg.det(tf.rot) - 1) < 1e-9
        assert 20.0 < ang < 45.0, (opt, ang)
