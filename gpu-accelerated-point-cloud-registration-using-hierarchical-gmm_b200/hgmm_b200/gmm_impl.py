"""Flat GMM EM with the reference's function surface (src/python/gmm_waymo/src/gmm_impl.py).

`train_gmm(X, max_iter, tol, means, covariances, weights, cov_type)` -> (inv_cov, means, weights,
covariances, log_ll) exactly as gmm_impl.py:118-145 returns it, but every E/M iteration is one fused
CUDA kernel in libhgmm instead of ~10 N x J NumPy/CuPy passes.  cov_type 'full' (the C++ fitter's
model, src/c++/gmm_fit/gmm_kernels.cu) is accepted as an extension.
"""
import contextlib
import time

import numpy as np

from .engine import Engine

eps = 1e-8     # gmm_impl.py:15

_default_engine = None


def default_engine():
    """process-wide Engine on cuda:0 (created on first use; raises if there is no GPU)."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def set_default_engine(engine):
    global _default_engine
    _default_engine = engine


def _host(a):
    if hasattr(a, "detach") and hasattr(a, "cpu"):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def init_gmm_params(X, k, cov_type="diag", rng=None):
    """gmm_impl.py:26-41: weights 1/k, covariances 0.1, means = k x d random *scalars* drawn from
    X.flatten() (sic).  `rng` (numpy Generator / RandomState) makes the draw reproducible; the
    reference uses the unseeded global RNG."""
    Xh = _host(X)
    weights = np.ones(k, dtype=np.float32) / k
    flat = Xh.reshape(-1)
    if rng is None:
        means = np.random.choice(flat, (k, Xh.shape[1]))
    else:
        means = rng.choice(flat, (k, Xh.shape[1]))
    if cov_type == "diag":
        covs = 0.1 * np.ones((k, Xh.shape[1]), dtype=np.float32)
    elif cov_type == "spherical":
        covs = 0.1 * np.ones((k,), dtype=np.float32)
    elif cov_type == "full":
        covs = np.tile(0.1 * np.eye(Xh.shape[1], dtype=np.float32), (k, 1, 1))
    else:
        raise ValueError("cov_type must be 'diag', 'spherical' or 'full'")
    return means.astype(np.float32), weights, covs


@contextlib.contextmanager
def timer(message):
    """gmm_impl.py:43-50 (the engine calls are synchronous on return, so no extra device sync)."""
    start = time.time()
    yield
    end = time.time()
    print('%s:  %f sec' % (message, end - start))


def train_gmm(X, max_iter, tol, means, covariances, weights, cov_type="diag", engine=None):
    """gmm_impl.py:118-145.  Returns (inv_cov, means, weights, covariances, log_ll).

    log_ll holds the mean log-likelihood of every executed iteration; like the reference, prints
    'Failed to converge...' when the tolerance was never met."""
    eng = engine or default_engine()
    eng.set_points(X)
    res = eng.fit_flat(_host(means), _host(covariances), _host(weights), cov_type=cov_type, max_iter=max_iter, tol=tol)
    lls = [float(v) for v in res["ll"]]
    converged = len(lls) >= 2 and abs(lls[-1] - lls[-2]) < tol
    if not converged:
        print('Failed to converge. Increase max-iter or tol.')
    eng._last_flat = (res["inv_cov"], res["means"], res["weights"], cov_type)
    return res["inv_cov"], res["means"], res["weights"], res["covs"], lls


def predict(X, inv_cov, means, weights, cov_type="diag", engine=None):
    """gmm_impl.py:147-155: argmax_j(log N_j(x) + log(pi_j + eps)).

    The model arguments are installed with a zero-iteration fit (covariance = 1/inv_cov^2), then
    the hard-assignment kernel runs over X."""
    eng = engine or default_engine()
    inv = _host(inv_cov).astype(np.float64)
    if cov_type == "full":
        raise ValueError("predict(): pass the model through Engine.fit_flat for full covariances")
    # a zero-iteration fit packs with the pre-loop rule inv_cov = 1/sqrt(cov) (gmm_impl.py:122),
    # so cov = 1/inv_cov^2 installs exactly the caller's inv_cov
    cov = 1.0 / (inv * inv)
    eng.set_points(X)
    eng.fit_flat(_host(means), cov, _host(weights), cov_type=cov_type, max_iter=0, tol=0.0, want_outputs=False)
    return eng.predict_flat().astype(np.int64)
