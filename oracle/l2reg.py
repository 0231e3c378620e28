"""L2-distance registration of two flat mixtures -- CPU oracle.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Restates src/python/gmmreg_gpu/{gmmreg,cost_functions,transforms,so}.py (citations relative to /root/reference/):
the mixtures are reduced to isotropic kernels of width sigma at the fitted means, the cost is minus the Gauss
transform between them, minimised over (quaternion, translation) by SciPy's BFGS.
Third-party arithmetic outside the reference checkout: `transformations.quaternion_matrix` (C. Gohlke's
transformations.py, no version pinned by the reference: README.md:254-263) -- restated from its published
definition; `scipy.optimize.minimize(method='BFGS')` -- used as is, exactly as the reference does, so the
optimiser's trajectory is SciPy's own in both.
"""
import math

import numpy as np

_EPS = np.finfo(float).eps * 4.0


def quaternion_matrix3(q):
    """3x3 block of transformations.quaternion_matrix; q = (w, x, y, z), normalised internally."""
    q = np.array(q, dtype=np.float64, copy=True)
    n = float(np.dot(q, q))
    if n < _EPS:
        return np.identity(3)
    q *= math.sqrt(2.0 / n)
    o = np.outer(q, q)
    return np.array([[1.0 - o[2, 2] - o[3, 3], o[1, 2] - o[3, 0], o[1, 3] + o[2, 0]],
                     [o[1, 2] + o[3, 0], 1.0 - o[1, 1] - o[3, 3], o[2, 3] - o[1, 0]],
                     [o[1, 3] - o[2, 0], o[2, 3] + o[1, 0], 1.0 - o[1, 1] - o[2, 2]]])


def diff_rot_from_quaternion(q):
    """so.py:4-59: d_rot[k] as the reference defines it (its diagonal entries are NOT the derivative of the normalised
    rotation -- e.g. d_rot[0,0,0] has the opposite sign convention in two of four entries -- they are reproduced as written)."""
    q = np.asarray(q, dtype=np.float64)
    rot = quaternion_matrix3(q)
    q2 = q * q
    z = q2.sum()
    z2 = z * z
    d = np.zeros((4, 3, 3))
    d[0, 0, 0] = 4 * q[0] * (q2[2] + q2[3]) / z2
    d[1, 0, 0] = 4 * q[1] * (q2[2] + q2[3]) / z2
    d[2, 0, 0] = -4 * q[2] * (q2[1] + q2[0]) / z2
    d[3, 0, 0] = -4 * q[3] * (q2[1] + q2[0]) / z2
    d[0, 1, 1] = 4 * q[0] * (q2[1] + q2[3]) / z2
    d[1, 1, 1] = -4 * q[1] * (q2[2] + q2[0]) / z2
    d[2, 1, 1] = 4 * q[2] * (q2[1] + q2[3]) / z2
    d[3, 1, 1] = -4 * q[3] * (q2[2] + q2[0]) / z2
    d[0, 2, 2] = 4 * q[0] * (q2[1] + q2[2]) / z2
    d[1, 2, 2] = -4 * q[1] * (q2[3] + q2[0]) / z2
    d[2, 2, 2] = -4 * q[2] * (q2[1] + q2[2]) / z2
    d[3, 2, 2] = 4 * q[3] * (q2[3] + q2[0]) / z2
    # off-diagonals: +-2 q_m / z - 2 q_k rot[a,b] / z2  (so.py:27-57); (sign, m) per [k][a,b]
    off = {(0, 1): [(-1, 3), (1, 2), (1, 1), (-1, 0)], (0, 2): [(1, 2), (1, 3), (1, 0), (1, 1)],
           (1, 0): [(1, 3), (1, 2), (1, 1), (1, 0)], (1, 2): [(-1, 1), (-1, 0), (1, 3), (1, 2)],
           (2, 0): [(-1, 2), (1, 3), (-1, 0), (1, 1)], (2, 1): [(1, 1), (1, 0), (1, 3), (1, 2)]}
    for (a, b), terms in off.items():
        for k, (sgn, m) in enumerate(terms):
            d[k, a, b] = sgn * 2 * q[m] / z - 2 * q[k] * rot[a, b] / z2
    return d


def gauss_transform(source, target, weights, h):
    """transforms.py:43-49: out[i] = sum_j weights[j] exp(-|target_i - source_j|^2 / h^2)"""
    d2 = ((target[:, None, :] - source[None, :, :]) ** 2).sum(axis=2)
    return np.exp(-d2 / (h * h)) @ weights


def compute_l2_dist(mu_source, phi_source, mu_target, phi_target, sigma):
    """cost_functions.py:29-40 -> (f, g [Js,3])"""
    z = np.power(2.0 * np.pi * sigma ** 2, mu_source.shape[1] * 0.5)
    h = np.sqrt(2.0) * sigma
    phi_j_e = gauss_transform(mu_target, mu_source, phi_target / z, h)
    phi_mu_j_e = np.stack([gauss_transform(mu_target, mu_source, w, h) for w in (phi_target * mu_target.T / z)]).T
    g = (phi_source * phi_j_e * mu_source.T - phi_source * phi_mu_j_e.T).T / (2.0 * sigma ** 2)
    return -np.dot(phi_source, phi_j_e), g


def rigid_cost(theta, mu_source, phi_source, mu_target, phi_target, sigma):
    """RigidCostFunction.__call__ (cost_functions.py:56-69) -> (f, grad[7]); theta = (qw,qx,qy,qz,tx,ty,tz)"""
    theta = np.asarray(theta, dtype=np.float64)
    rot = quaternion_matrix3(theta[:4])
    t_mu = np.dot(mu_source, rot.T) + theta[4:7]
    f, g = compute_l2_dist(t_mu, phi_source, mu_target, phi_target, sigma)
    d_rot = diff_rot_from_quaternion(theta[:4])
    gtm0 = np.dot(g.T, mu_source)
    grad = np.concatenate([(gtm0 * d_rot).sum(axis=(1, 2)), g.sum(axis=0)])
    return f, grad


def estimate_sigma(data):
    """gmmreg.py:48-52"""
    n, d = data.shape
    dh = data - data.mean(axis=0)
    return float(np.power(np.linalg.det(np.dot(dh.T, dh) / (n - 1)), 1.0 / (2.0 * d)))


def registration(mu_source_fn, mu_target, phi_target, sigma, delta=0.9, maxiter=1, tol=1.0e-3, opt_maxiter=10, opt_tol=1.0e-5):
    """L2DistRegistration.registration (gmmreg.py:62-121) with the feature generator abstracted:
    `mu_source_fn()` returns (mu_source, phi_source) (the reference re-fits the source every outer iteration);
    weights are scaled by 1e3 here as the reference does (:75,:88).  -> (rot, t, theta, f)"""
    from scipy.optimize import minimize
    x = np.zeros(7)
    x[0] = 1.0
    phi_t = phi_target * 1e3
    f_prev = None
    res = None
    for _ in range(maxiter):
        mu_s, phi_s = mu_source_fn()
        phi_s = phi_s * 1e3
        res = minimize(rigid_cost, x, args=(mu_s, phi_s, mu_target, phi_t, sigma), method="BFGS", jac=True, tol=opt_tol,
                       options={"maxiter": opt_maxiter})
        sigma *= delta
        if f_prev is not None and abs(res.fun - f_prev) < tol:
            break
        f_prev = res.fun
        x = res.x
    return quaternion_matrix3(res.x[:4]), res.x[4:7].copy(), res.x.copy(), float(res.fun)
