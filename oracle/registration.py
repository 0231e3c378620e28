"""Tree registration -- CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates src/python/hgmm/hgmm_gpu.py:550-577 (gmmTreeRegESTep), :620-664 (twist helpers),
:729-768 (GMMTree.maximization_step / registration); the north-star weighted-Procrustes solve;
and the McAdams 3x3 SVD of src/c++/common/svd3.h.  float64 throughout.
"""
import numpy as np

from .hgmm_tree import EPS, N_NODE, child, children_gamma, complexity, n_total

F32_EPS = float(np.finfo(np.float32).eps)     # hgmm_gpu.py:741


def reg_e_step(X, pi, mu, cov, max_level, lambda_c):
    """gmmTreeRegESTep (hgmm_gpu.py:550-577): greedy root->leaf descent per point.

    At each level: gamma over the 8 children of the current node, normalised (zeros if den<=EPS,
    :563-567); searchID = j0 + argmax; stop BEFORE accumulating if complexity(cov[searchID]) <=
    lambda_c (:572-573); else accumulate (gamma_max, gamma_max x, gamma_max x x^T) into that node
    unless gamma_max < EPS (:457-459).  Vectorised over points with an `alive` mask.
    Returns (M0 [nt], M1 [nt,3], M2 [nt,3,3])."""
    X = np.asarray(X, dtype=np.float64)
    nt = n_total(max_level)
    cplx = complexity(cov)
    M0 = np.zeros(nt)
    M1 = np.zeros((nt, 3))
    M2 = np.zeros((nt, 3, 3))
    search = -np.ones(X.shape[0], dtype=np.int64)
    alive = np.ones(X.shape[0], dtype=bool)
    for _ in range(max_level):
        if not alive.any():
            break
        ia = np.nonzero(alive)[0]
        g, j0 = children_gamma(X[ia], pi, mu, cov, search[ia])
        den = g.sum(axis=1)
        gn = np.where((den > EPS)[:, None], g / np.where(den > EPS, den, 1.0)[:, None], 0.0)
        mx = gn.argmax(axis=1)
        sid = j0 + mx
        search[ia] = sid
        stop = cplx[sid] <= lambda_c
        alive[ia[stop]] = False
        gm = gn[np.arange(len(ia)), mx]
        use = (~stop) & (gm >= EPS)
        w = gm[use]
        xs = X[ia[use]]
        M0 += np.bincount(sid[use], weights=w, minlength=nt)
        for a in range(3):
            M1[:, a] += np.bincount(sid[use], weights=w * xs[:, a], minlength=nt)
            for b in range(3):
                M2[:, a, b] += np.bincount(sid[use], weights=w * xs[:, a] * xs[:, b], minlength=nt)
    return M0, M1, M2


def skew(x):
    """hgmm_gpu.py:620-631."""
    return np.array([[0.0, -x[2], x[1]], [x[2], 0.0, -x[0]], [-x[1], x[0], 0.0]])


def twist_trans(tw):
    """hgmm_gpu.py:646-664 (non-linear branch): Rodrigues rotation of tw[:3], translation tw[3:]."""
    th = np.linalg.norm(tw[:3])
    if th == 0.0:
        return np.identity(3), tw[3:]
    n = tw[:3] / th
    return (np.cos(th) * np.identity(3) + (1.0 - np.cos(th)) * np.outer(n, n) + np.sin(th) * skew(n)), tw[3:]


def twist_mul(tw, rot, t):
    """hgmm_gpu.py:634-644: (R, t) <- (dR R, dR t + dt)."""
    tr, tt = twist_trans(tw)
    return tr @ rot, t @ tr.T + tt


def reg_m_step_lstsq(M0, M1, pi, mu, cov, rot, t):
    """GMMTree.maximization_step (hgmm_gpu.py:729-752), literally: per node with M0 >= f32-eps,
    (lam, n) = eigh(cov_i); s = M1/M0; n *= sqrt(M0/lam); b = n^T mu - n^T s; A = [s x n_k | n_k];
    x, q = lstsq(A, b); (R,t) <- twist_mul(x, R, t).  Returns (R, t, q, x); q is the residual
    sum of squares (numpy returns it as shape (1,) or empty when rank-deficient)."""
    n = len(pi)
    A = np.zeros((3 * n, 6))
    b = np.zeros(3 * n)
    for i in range(n):
        if M0[i] < F32_EPS:
            continue
        lam, nn = np.linalg.eigh(cov[i])
        s = M1[i] / M0[i]
        nn = nn * np.sqrt(M0[i] / lam)
        sl = slice(3 * i, 3 * i + 3)
        b[sl] = nn.T @ mu[i] - nn.T @ s
        A[sl, :3] = np.cross(s, nn.T)
        A[sl, 3:] = nn.T
    x, q, _, _ = np.linalg.lstsq(A, b, rcond=-1)
    r2, t2 = twist_mul(x, rot, t)
    return r2, t2, q, x


def reg_normal_equations(M0, M1, mu, cov):
    """The eigh-free form of the same least squares: since
    sum_k n_k n_k^T / lam_k = cov^-1, the system is the normal equations of
    min sum_i M0_i || J_i x - (mu_i - s_i) ||^2_{cov_i^-1},  J_i = [-[s_i]x | I3].
    Returns (H [6,6], g [6], c) with x = H^-1 g and residual q = c - g^T x."""
    H = np.zeros((6, 6))
    g = np.zeros(6)
    c = 0.0
    for i in np.nonzero(M0 >= F32_EPS)[0]:
        s = M1[i] / M0[i]
        P = np.linalg.inv(cov[i])
        J = np.hstack([-skew(s), np.identity(3)])
        r = mu[i] - s
        H += M0[i] * J.T @ P @ J
        g += M0[i] * J.T @ P @ r
        c += M0[i] * r @ P @ r
    return H, g, c


def reg_m_step_procrustes(M0, M1, mu, rot, t):
    """North-star solver: weighted Procrustes between the moment centroids s_i = M1_i/M0_i
    (where the transformed target mass sits) and the node means mu_i, weights M0_i;
    R = V diag(1,1,det) U^T of W = sum w (s-sbar)(mu-mubar)^T (reflection-fixed, unlike
    icp_kernel.cu:718-729), dt = mubar - dR sbar; composed onto (rot, t) like twist_mul."""
    use = M0 >= F32_EPS
    w = M0[use]
    s = M1[use] / w[:, None]
    m = mu[use]
    sb = (w[:, None] * s).sum(0) / w.sum()
    mb = (w[:, None] * m).sum(0) / w.sum()
    W = ((s - sb) * w[:, None]).T @ (m - mb)
    U, S, Vt = np.linalg.svd(W)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    dR = Vt.T @ D @ U.T
    dt = mb - dR @ sb
    q = float((w * ((s @ dR.T + dt - m) ** 2).sum(1)).sum())
    return dR @ rot, dR @ t + dt, q


def registration(target, pi, mu, cov, max_level, lambda_c=0.01, maxiter=20, tol=1.0e-4,
                 solver="twist_lstsq", rot=None, t=None):
    """GMMTree.registration (hgmm_gpu.py:754-768): iterate transform(target) -> E -> M until
    |q - q_prev| < tol; returns the INVERSE transform (R^T, -R^T t) (:768), the last q and the
    iteration count."""
    Y = np.asarray(target, dtype=np.float64)
    rot = np.identity(3) if rot is None else np.array(rot, dtype=np.float64)
    t = np.zeros(3) if t is None else np.array(t, dtype=np.float64)
    q_prev = None
    it = 0
    for it in range(1, maxiter + 1):
        ty = Y @ rot.T + t
        M0, M1, _ = reg_e_step(ty, pi, mu, cov, max_level, lambda_c)
        if solver == "twist_lstsq":
            rot, t, q, _ = reg_m_step_lstsq(M0, M1, pi, mu, cov, rot, t)
            q = float(q[0]) if np.size(q) else float("nan")
        else:
            rot, t, q = reg_m_step_procrustes(M0, M1, mu, rot, t)
        if q_prev is not None and abs(q - q_prev) < tol:
            break
        q_prev = q
    return rot.T, -rot.T @ t, q, it


# --------------------------------------------------------------------------------------
# Flat-mixture registration (BASELINE configs[3]).  The reference's entry point for it,
# GMMRegistration::pointCloudRegisterGPU (src/c++/gmm_registration/gmm_reg.cu:54-56), is an empty
# stub, so there is nothing of the reference's to pin this against: PARITY UNPINNED BY THE REFERENCE.
# What is restated is the north star's definition -- the flat E-step (gmm_kernels.cu:278-302
# responsibilities, as oracle/flat_gmm.py::cpp_em_iteration evaluates them) used as point-to-mixture
# correspondence, feeding the same two solves as the tree registration above.
# --------------------------------------------------------------------------------------
def flat_reg_moments(Y, weights, mu, cov):
    """responsibilities of the (already transformed) target over all J components of a full-covariance
    mixture -> (S0 [J] = sum_i gamma_ij, S1 [J,3] = sum_i gamma_ij y_i)."""
    Y = np.asarray(Y, dtype=np.float64)
    d = Y[:, None, :] - mu[None, :, :]
    inv = np.linalg.inv(cov)
    maha = np.einsum("nja,jab,njb->nj", d, inv, d)
    with np.errstate(divide="ignore"):
        a = np.log(weights)[None, :] - 0.5 * (3.0 * np.log(2.0 * np.pi) + np.log(np.linalg.det(cov))[None, :] + maha)
    m = a.max(axis=1, keepdims=True)
    e = np.exp(a - m)
    g = e / e.sum(axis=1, keepdims=True)
    return g.sum(axis=0), g.T @ Y


def flat_registration(target, weights, mu, cov, maxiter=20, tol=1.0e-4, solver="procrustes", rot=None, t=None):
    """the loop of `registration` above with the flat E-step; returns the INVERSE transform (R^T, -R^T t),
    the last q and the iteration count (same conventions as GMMTree.registration, hgmm_gpu.py:754-768)."""
    Y = np.asarray(target, dtype=np.float64)
    weights, mu, cov = (np.asarray(v, dtype=np.float64) for v in (weights, mu, cov))
    rot = np.identity(3) if rot is None else np.array(rot, dtype=np.float64)
    t = np.zeros(3) if t is None else np.array(t, dtype=np.float64)
    q_prev = None
    it = 0
    for it in range(1, maxiter + 1):
        S0, S1 = flat_reg_moments(Y @ rot.T + t, weights, mu, cov)
        if solver == "twist_lstsq":
            H, g, c = reg_normal_equations(S0, S1, mu, cov)
            x = np.linalg.solve(H, g)
            q = float(c - g @ x)
            rot, t = twist_mul(x, rot, t)
        else:
            rot, t, q = reg_m_step_procrustes(S0, S1, mu, rot, t)
        if q_prev is not None and abs(q - q_prev) < tol:
            break
        q_prev = q
    return rot.T, -rot.T @ t, q, it


# --------------------------------------------------------------------------------------
# McAdams / Selle / Tamstorf / Teran / Sifakis 3x3 SVD as arranged in common/svd3.h
# --------------------------------------------------------------------------------------
_GAMMA = 5.828427124      # svd3.h: 3 + 2*sqrt(2)
_CSTAR = 0.923879532      # cos(pi/8)
_SSTAR = 0.3826834323     # sin(pi/8)


def _approx_givens(a11, a12, a22):
    """svd3.h approximateGivensQuaternion: (ch, sh) of the Jacobi rotation, with the exact rsqrt."""
    ch = 2.0 * (a11 - a22)
    sh = a12
    b = _GAMMA * sh * sh < ch * ch
    w = 1.0 / np.sqrt(ch * ch + sh * sh) if (ch * ch + sh * sh) > 0 else 0.0
    return (w * ch, w * sh) if b else (_CSTAR, _SSTAR)


def svd3_mcadams(A, sweeps=4):
    """svd (svd3.h:355-401): Jacobi eigen-analysis of A^T A (4 sweeps, :231-240) -> V; B = A V;
    sort columns by decreasing norm (negating to keep det V = +1); Givens QR of B -> U, S.
    Returns (U, S_diag_matrix, V) with A ~= U S V^T.  Plain float64 restatement used to check the
    device warp-SVD; np.linalg.svd is the independent cross-check."""
    A = np.asarray(A, dtype=np.float64)
    S = A.T @ A
    V = np.identity(3)
    for _ in range(sweeps):
        for (p, q) in ((0, 1), (1, 2), (0, 2)):
            ch, sh = _approx_givens(S[p, p], S[p, q], S[q, q])
            c = ch * ch - sh * sh
            s = 2.0 * ch * sh
            n = ch * ch + sh * sh
            c, s = c / n, s / n
            G = np.identity(3)
            G[p, p] = c
            G[q, q] = c
            G[p, q] = -s
            G[q, p] = s
            S = G.T @ S @ G
            V = V @ G
    B = A @ V
    norms = (B * B).sum(axis=0)
    order = [0, 1, 2]
    for (i, j) in ((0, 1), (0, 2), (1, 2)):       # conditional swaps, svd3.h sortSingularValues
        if norms[order[i]] < norms[order[j]]:
            order[i], order[j] = order[j], order[i]
            B[:, order[i]] *= 1.0
    Bs = B[:, order].copy()
    Vs = V[:, order].copy()
    if np.linalg.det(Vs) < 0:                      # the reference negates a swapped column
        Bs[:, 2] *= -1.0
        Vs[:, 2] *= -1.0
    Q, R = np.linalg.qr(Bs)
    sg = np.sign(np.diag(R))
    sg[sg == 0] = 1.0
    return Q * sg[None, :], np.diag(np.diag(R) * sg), Vs
