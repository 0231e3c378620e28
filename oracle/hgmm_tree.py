"""Hierarchical (8-ary tree) GMM build -- CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates src/python/hgmm/hgmm_cupy_cpu_working.py (the semantic authority for the degenerate
cases) cross-checked with src/python/hgmm/hgmm_gpu.py; vectorised over
points (the reference loops per point in pure Python) and evaluated in float64.
Citations are relative to /root/reference/src/python/hgmm/.
"""
import numpy as np

N_NODE = 8           # hgmm_gpu.py:30
EPS = 1.0e-15        # hgmm_gpu.py:29 / hgmm_cupy_cpu_working.py:30


def child(j):
    """first child of node j; j=-1 is the root (hgmm_gpu.py:84-89)."""
    return (j + 1) * N_NODE


def level(l):
    """offset of tree level l in the node array (hgmm_gpu.py:91-92): 0, 8, 72, 584, 4680, 37448."""
    return N_NODE * (N_NODE ** l - 1) // (N_NODE - 1)


def n_total(max_level):
    """hgmm_gpu.py:467."""
    return level(max_level)


def reference_init_indices(max_level, flavor="gpu"):
    """The seeds the reference draws: hgmm_gpu.py:469-470 (np.random.seed(72); randint(nTotal, size=nTotal));
    the CPU file seeds with nTotal instead (hgmm_cupy_cpu_working.py:124-125, CuPy RNG in the original)."""
    nt = n_total(max_level)
    rs = np.random.RandomState(72 if flavor == "gpu" else nt)
    return rs.randint(nt, size=nt)


def gaussian_pdf_nodes(X, mu, cov):
    """N(x_i; mu_i[k], cov_i[k]) for per-point gathered nodes; 0 where det < EPS
    (hgmm_cupy_cpu_working.py:62-70, hgmm_gpu.py:284-312).  X [N,3], mu [N,K,3], cov [N,K,3,3]."""
    det = np.linalg.det(cov)
    ok = det >= EPS
    safe = np.where(ok[..., None, None], cov, np.eye(3))
    inv = np.linalg.inv(safe)
    d = X[:, None, :] - mu
    maha = np.einsum("nka,nkab,nkb->nk", d, inv, d)
    c = 1.0 / (np.sqrt(np.where(ok, det, 1.0)) * (2.0 * np.pi) ** 1.5)
    return np.where(ok, c * np.exp(-0.5 * maha), 0.0)


def _node_tables(mu, cov):
    det = np.linalg.det(cov)
    ok = det >= EPS
    inv = np.linalg.inv(np.where(ok[:, None, None], cov, np.eye(3)))
    c = 1.0 / (np.sqrt(np.where(ok, det, 1.0)) * (2.0 * np.pi) ** 1.5)
    return ok, inv, c


def children_gamma(X, pi, mu, cov, parent):
    """Unnormalised gamma_k = pi_k N(x; k) over the 8 children of `parent` (hgmm_gpu.py:393-402).
    Returns (gamma_raw [N,8], j0 [N])."""
    ok, inv, c = _node_tables(mu, cov)
    j0 = child(parent)
    idx = j0[:, None] + np.arange(N_NODE)[None, :]
    d = X[:, None, :] - mu[idx]
    maha = np.einsum("nka,nkab,nkb->nk", d, inv[idx], d)
    pdf = np.where(ok[idx], c[idx] * np.exp(-0.5 * maha), 0.0)
    return pi[idx] * pdf, j0


def tree_e_step(X, pi, mu, cov, parent, nt, skip=None):
    """gmmTreeEStep (hgmm_cupy_cpu_working.py:162-191) + accumulate (:99-106).

    gamma = gamma/den if den > EPS else 0 (:174-178); a child is accumulated only if its
    normalised gamma >= EPS (:100-101); currentIdx = j0 + argmax(gamma) (first max, 0 if all zero).
    Returns (M0 [nt], M1 [nt,3], M2 [nt,3,3], current [N], den [N])."""
    g, j0 = children_gamma(X, pi, mu, cov, parent)
    if skip is not None:                       # adaptive build: points under a terminal node take no part (child 0, no moments, no q)
        g = np.where(skip[:, None], 0.0, g)
    den = g.sum(axis=1)
    gn = np.where((den > EPS)[:, None], g / np.where(den > EPS, den, 1.0)[:, None], 0.0)
    current = j0 + gn.argmax(axis=1)
    acc = np.where(gn < EPS, 0.0, gn)
    idx = (j0[:, None] + np.arange(N_NODE)[None, :]).ravel()
    M0 = np.bincount(idx, weights=acc.ravel(), minlength=nt)
    M1 = np.zeros((nt, 3))
    M2 = np.zeros((nt, 3, 3))
    for a in range(3):
        M1[:, a] = np.bincount(idx, weights=(acc * X[:, a:a + 1]).ravel(), minlength=nt)
        for b in range(3):
            M2[:, a, b] = np.bincount(idx, weights=(acc * (X[:, a] * X[:, b])[:, None]).ravel(), minlength=nt)
    return M0, M1, M2, current, den


def ml_estimator(M0, M1, M2, n_points, ld):
    """mlEstimator (hgmm_cupy_cpu_working.py:109-119): blank node (pi=0, mu=0, cov=I) if M0 < ld,
    else pi = M0/N_points, mu = M1/M0, cov = M2/M0 - mu mu^T.  Vectorised over nodes."""
    live = M0 >= ld
    m0 = np.where(live, M0, 1.0)
    pi = np.where(live, M0 / n_points, 0.0)
    mu = np.where(live[:, None], M1 / m0[:, None], 0.0)
    cov = M2 / m0[:, None, None] - mu[:, :, None] * mu[:, None, :]
    cov = np.where(live[:, None, None], cov, np.eye(3))
    return pi, mu, cov


def level_log_likelihood(X, pi, mu, cov, lb, le, chunk=4096):
    """logLikelihoodValue (hgmm_cupy_cpu_working.py:72-85): sum_i log max(sum_{j in [lb,le), pi_j>=EPS} pi_j N(x_i;j), EPS)."""
    ok, inv, c = _node_tables(mu[lb:le], cov[lb:le])
    w = np.where(ok & (pi[lb:le] >= EPS), pi[lb:le] * c, 0.0)
    q = 0.0
    for s in range(0, X.shape[0], chunk):
        d = X[s:s + chunk, None, :] - mu[None, lb:le, :]
        maha = np.einsum("nka,kab,nkb->nk", d, inv, d)
        p = (w[None, :] * np.exp(-0.5 * maha)).sum(axis=1)
        q += np.log(np.maximum(p, EPS)).sum()
    return q


def build_gmm_tree(points, max_level, ls, ld, init_means, sig2=0.004, ll_mode="level",
                   max_iters_per_level=10000, return_trace=False, prune_lambda_c=0.0, prune_min_points=0.0):
    """buildGMMTree (hgmm_gpu.py:466-548, hgmm_cupy_cpu_working.py:122-160).

    init: every node pi=1/8, mu=init_means[i], cov=sig2*I (hgmm_gpu.py:487-490);
    per level: repeat E, M, q=log-lik until |q-prevQ| < ls with prevQ=0 at level start (:519-538);
    then parent <- current (:540).
    ll_mode "level": the reference's scan over the whole level with the NEW parameters;
    ll_mode "estep": q = sum_i log max(den_i, EPS) from the E-step's own 8-sibling normaliser
                     (the engine's fast mode; lags the reference by one iteration).
    Returns (pi [nt], mu [nt,3], cov [nt,3,3], current [N]) (+ per-level iteration counts / q trace)."""
    X = np.asarray(points, dtype=np.float64)
    N = X.shape[0]
    nt = n_total(max_level)
    pi = np.full(nt, 1.0 / N_NODE)
    mu = np.array(init_means, dtype=np.float64).reshape(nt, 3).copy()
    cov = np.tile(np.eye(3) * sig2, (nt, 1, 1))
    parent = -np.ones(N, dtype=np.int64)
    current = np.zeros(N, dtype=np.int64)
    iters, trace = [], []
    terminal = np.zeros(nt, dtype=bool)        # adaptive build (include/hgmm.h hgmm_tree_config.prune_*): not in the reference
    for l in range(max_level):
        lb, le = level(l), level(l + 1)
        prev_q = 0.0
        it = 0
        skip = terminal[np.maximum(parent, 0)] & (parent >= 0) if (prune_lambda_c > 0 or prune_min_points > 0) else None
        while True:
            M0, M1, M2, current, den = tree_e_step(X, pi, mu, cov, parent, nt, skip)
            npi, nmu, ncov = ml_estimator(M0[lb:le], M1[lb:le], M2[lb:le], N, ld)
            pi[lb:le], mu[lb:le], cov[lb:le] = npi, nmu, ncov
            if ll_mode == "level":
                q = level_log_likelihood(X, pi, mu, cov, lb, le)
            elif skip is not None:
                q = float(np.log(np.maximum(den[~skip], EPS)).sum())
            else:
                q = float(np.log(np.maximum(den, EPS)).sum())
            it += 1
            trace.append((l, it, q))
            if abs(q - prev_q) < ls or it >= max_iters_per_level:
                break
            prev_q = q
        iters.append(it)
        parent = current.copy()
        if prune_lambda_c > 0 or prune_min_points > 0:
            with np.errstate(invalid="ignore", divide="ignore"):
                cx = complexity(cov[lb:le])
            terminal[lb:le] = (pi[lb:le] <= 0) | ((prune_min_points > 0) & (pi[lb:le] * N < prune_min_points)) | \
                              ((prune_lambda_c > 0) & (cx <= prune_lambda_c))
    if return_trace:
        return pi, mu, cov, current, iters, trace
    return pi, mu, cov, current


def complexity(cov):
    """complexity (hgmm_gpu.py:78-82): smallest eigenvalue / trace of a node covariance. Vectorised."""
    lam = np.linalg.eigvalsh(np.asarray(cov, dtype=np.float64))
    return lam[..., 0] / lam.sum(axis=-1)
