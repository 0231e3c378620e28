"""Flat (non-hierarchical) GMM EM -- CPU oracle.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Two reference variants are restated here; citations are relative to /root/reference/.

``py``  : src/python/gmm_waymo/src/gmm_impl.py  (diag / spherical covariance, NumPy-or-CuPy)
``cpp`` : src/c++/gmm_fit/gmm_kernels.cu         (full 3x3 covariance, CUDA)

Both are written directly from the maths (distances about the mean, float64 by default)
rather than in the reference's expanded-GEMM float32 form; `make_golden.py` checks that they
agree with the unmodified reference run in float32 to ~1e-5.
"""
import numpy as np

EPS = 1e-8          # gmm_impl.py:15
REG_COVAR = 1e-6    # gmm_impl.py:81 (estimate_covariance default) and :134 (inside the sqrt)


# --------------------------------------------------------------------------------------
# ``py`` variant
# --------------------------------------------------------------------------------------
def py_inv_cov_from_cov(cov, first=False):
    """gmm_impl.py:122 (before the loop: 1/sqrt(cov)) and :134 (inside: 1/(sqrt(cov+1e-6)+eps))."""
    cov = np.asarray(cov)
    if first:
        return 1.0 / np.sqrt(cov)
    return 1.0 / (np.sqrt(cov + REG_COVAR) + EPS)


def py_log_prob(X, inv_cov, means, cov_type="diag"):
    """log N(x_i; mu_j, diag) as the reference defines it (gmm_impl.py:53-78).

    `inv_cov` is 1/std ([J,3] for diag, [J] for spherical).  Reference quirks kept:
    log-det term is sum(log(inv_cov + eps)) (diag, :70) or 3*log(inv_cov + eps) (spherical, :57).
    """
    X = np.asarray(X)
    inv_cov = np.asarray(inv_cov, dtype=X.dtype)
    means = np.asarray(means, dtype=X.dtype)
    nfeat = X.shape[1]
    d = X[:, None, :] - means[None, :, :]                      # [N,J,3]
    if cov_type == "diag":
        prec = inv_cov ** 2                                     # [J,3]
        maha = np.einsum("njk,jk->nj", d * d, prec)
        log_det = np.sum(np.log(inv_cov + EPS), axis=1)
    elif cov_type == "spherical":
        prec = inv_cov ** 2                                     # [J]
        maha = np.einsum("njk,njk->nj", d, d) * prec[None, :]
        log_det = nfeat * np.log(inv_cov + EPS)
    else:
        raise ValueError(cov_type)
    return -0.5 * (nfeat * np.log(2.0 * np.pi) + maha) + log_det[None, :]


def py_e_step(X, inv_cov, means, weights, cov_type="diag"):
    """gmm_impl.py:105-116.  No max-shift in the reference; `+eps` inside both logs."""
    wlp = py_log_prob(X, inv_cov, means, cov_type) + np.log(np.asarray(weights) + EPS)[None, :]
    norm = np.log(np.sum(np.exp(wlp), axis=1) + EPS)
    return norm.mean(), wlp - norm[:, None], norm


def py_m_step(X, resp, cov_type="diag"):
    """gmm_impl.py:81-103: nk = sum(resp)+eps; mu = resp^T X / nk;
    cov = E[x^2] - 2 mu E[x] + mu^2 + 1e-6 (diag); spherical = mean over the 3 dims."""
    nk = resp.sum(axis=0) + EPS
    sx = resp.T @ X
    means = sx / nk[:, None]
    cov = (resp.T @ (X * X)) / nk[:, None] - 2.0 * means * sx / nk[:, None] + means ** 2 + REG_COVAR
    if cov_type == "spherical":
        cov = cov.mean(axis=1)
    return nk / X.shape[0], means, cov


def py_train_gmm(X, max_iter, tol, means, covariances, weights, cov_type="diag", dtype=np.float64):
    """gmm_impl.py:118-145.  Returns (inv_cov, means, weights, covariances, log_ll list)."""
    X = np.asarray(X, dtype=dtype)
    means = np.asarray(means, dtype=dtype)
    covariances = np.asarray(covariances, dtype=dtype)
    weights = np.asarray(weights, dtype=dtype)
    inv_cov = py_inv_cov_from_cov(covariances, first=True)
    lower = -np.inf
    lls = []
    for _ in range(max_iter):
        prev = lower
        ll, log_resp, _ = py_e_step(X, inv_cov, means, weights, cov_type)
        lls.append(ll)
        weights, means, covariances = py_m_step(X, np.exp(log_resp), cov_type)
        inv_cov = py_inv_cov_from_cov(covariances)
        lower = ll
        if abs(lower - prev) < tol:
            break
    return inv_cov, means, weights, covariances, lls


# --------------------------------------------------------------------------------------
# ``py_old`` variant: src/python/gmmreg_gpu/gmm_impl.py (the copy the L2 registration path fits with).
# Diagonal only.  Differences from the gmm_waymo copy: log(weights) without eps (:58), nk without eps and
# weights = nk/N (:47,52), means and E[x^2] divided by (nk + eps) (:48-49), covariance = clip(E[x^2] - mu^2, 0)
# with no 1e-6 floor (:51), inv_cov = 1/(sqrt(cov) + eps) (:71).
# --------------------------------------------------------------------------------------
def py_old_train_gmm(X, max_iter, tol, means, covariances, weights, dtype=np.float64):
    """gmmreg_gpu/gmm_impl.py:62-83.  Returns (inv_cov, means, weights, covariances, log_ll list)."""
    X = np.asarray(X, dtype=dtype)
    means = np.asarray(means, dtype=dtype)
    covariances = np.asarray(covariances, dtype=dtype)
    weights = np.asarray(weights, dtype=dtype)
    inv_cov = 1.0 / np.sqrt(covariances)
    lower = -np.inf
    lls = []
    for _ in range(max_iter):
        prev = lower
        with np.errstate(divide="ignore"):
            wlp = py_log_prob(X, inv_cov, means, "diag") + np.log(weights)[None, :]
        norm = np.log(np.sum(np.exp(wlp), axis=1) + EPS)
        ll = norm.mean()
        lls.append(ll)
        resp = np.exp(wlp - norm[:, None])
        nk = resp.sum(axis=0)
        means = (resp.T @ X) / (nk[:, None] + EPS)
        x2 = (resp.T @ (X * X)) / (nk[:, None] + EPS)
        covariances = np.clip(x2 - means ** 2, 0.0, None)
        weights = nk / X.shape[0]
        inv_cov = 1.0 / (np.sqrt(covariances) + EPS)
        lower = ll
        if abs(lower - prev) < tol:
            break
    return inv_cov, means, weights, covariances, lls


def py_predict(X, inv_cov, means, weights, cov_type="diag"):
    """gmm_impl.py:147-155: argmax_j(log_prob + log(pi + eps))."""
    X = np.asarray(X)
    lp = py_log_prob(X, inv_cov, means, cov_type) + np.log(np.asarray(weights) + EPS)[None, :]
    return lp.argmax(axis=1)


# --------------------------------------------------------------------------------------
# ``cpp`` variant (full covariance)
# --------------------------------------------------------------------------------------
def cpp_log_gauss(X, mu, cov, sigma_bug=False):
    """gmm_kernels.cu:96-126: -0.5*(3 log 2pi + log|S| + d^T M d).

    The reference computes glm::inverse(S) (:97) but then multiplies by S itself (:103);
    `sigma_bug=True` reproduces that (M = S), the default is the intended M = S^-1."""
    d = X[:, None, :] - mu[None, :, :]
    M = cov if sigma_bug else np.linalg.inv(cov)
    maha = np.einsum("nja,jab,njb->nj", d, M, d)
    logdet = np.log(np.linalg.det(cov))
    return -0.5 * (3.0 * np.log(2.0 * np.pi) + logdet[None, :] + maha)


def cpp_em_iteration(X, logpi, mu, cov, sigma_bug=False):
    """One iteration of the loop at gmm_kernels.cu:455-465.

    E (:278-302): prob_ij = logN_ij - log sum_k exp(logpi_k + logN_ik)   (no logpi_j term!)
    M (:304-350): T_j = logsumexp_i prob_ij (:135-154); logpi_j += T_j - log sum_k exp(T_k+logpi_k)
                  (:156-179); mu_j = sum_i x_i e^{prob_ij} / e^{T_j} (:181-192);
                  S_j  = sum_i (x_i-mu_j)(x_i-mu_j)^T e^{prob_ij} / e^{T_j} with the NEW mu (:194-210).
    Evaluated here with a max-shifted log-sum-exp (the reference has none, :290-294), which is
    identical wherever the reference does not over/underflow.
    Returns (logpi, mu, cov, data log-likelihood sum_i log p(x_i) under the INPUT parameters).
    """
    logn = cpp_log_gauss(X, mu, cov, sigma_bug)
    a = logn + logpi[None, :]
    m = a.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(a - m).sum(axis=1, keepdims=True)))[:, 0]
    prob = logn - lse[:, None]
    pm = prob.max(axis=0)
    w = np.exp(prob - pm[None, :])                 # e^{prob_ij} / e^{pm_j}
    T = pm + np.log(w.sum(axis=0))
    z = T + logpi
    zm = z.max()
    new_logpi = z - (zm + np.log(np.exp(z - zm).sum()))
    wn = w / w.sum(axis=0, keepdims=True)          # e^{prob_ij} / e^{T_j}
    new_mu = wn.T @ X
    d = X[:, None, :] - new_mu[None, :, :]
    new_cov = np.einsum("nj,nja,njb->jab", wn, d, d)
    return new_logpi, new_mu, new_cov, lse.sum()


def cpp_fit(X, mu0, iterations, sigma0_sq=1.0, sigma_bug=False, dtype=np.float64):
    """GMM::solve (gmm_kernels.cu:371-504) with the random init passed in:
    S = sigma0_sq*I (:397-402, reference uses I), logpi = log(1/J) (:393-395), `iterations` EM steps.
    Returns (weights, mu, cov, [log-likelihood before each M step])."""
    X = np.asarray(X, dtype=dtype)
    mu = np.array(mu0, dtype=dtype)
    J = mu.shape[0]
    cov = np.tile(np.eye(3, dtype=dtype) * sigma0_sq, (J, 1, 1))
    logpi = np.full(J, -np.log(J), dtype=dtype)
    lls = []
    for _ in range(iterations):
        logpi, mu, cov, ll = cpp_em_iteration(X, logpi, mu, cov, sigma_bug)
        lls.append(ll)
    return np.exp(logpi), mu, cov, lls


def rel_fro(a, b):
    """relative Frobenius distance ||a-b||_F / ||b||_F (the parity metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
