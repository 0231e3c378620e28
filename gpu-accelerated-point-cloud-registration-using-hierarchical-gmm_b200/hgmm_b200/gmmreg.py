"""L2-distance (GMMReg) registration with the reference's surface (src/python/gmmreg_gpu/gmmreg.py,
cost_functions.py, gmm.py, gmm_impl.py).

    registration_gmmreg(source, target, tf_type_name='rigid', callbacks=[], **kargs) -> RigidTransformation
    RigidGMMReg, L2DistRegistration, RigidCostFunction, GMM_GPU (the older diagonal fitter of that directory)

What runs where: both mixture fits are the fused E+M CUDA sweep (flavour HGMM_FLAVOR_PY_OLD reproduces
gmmreg_gpu/gmm_impl.py's update rules), the J x J Gauss-transform cost and its quaternion gradient are one float64
CUDA kernel, and the minimisation is either
    optimizer='scipy'  : scipy.optimize.minimize(method='BFGS', jac=True), as gmmreg.py:101-107 -- the reference's own
                         third-party optimiser driving the device cost function (one kernel launch per evaluation); or
    optimizer='device' : the whole BFGS + strong-Wolfe line search in a single kernel launch (hgmm_l2_optimize).
There is no CPU implementation of the cost here; without libhgmm / a GPU every entry point raises.
"""
import math
import time

import numpy as np

from . import gmm_impl as _gi
from ._lib import FLAVOR_PY_OLD
from .hgmm import RigidTransformation

_EPS4 = np.finfo(float).eps * 4.0


def _cv(x):
    return np.asarray(x.points if hasattr(x, "points") else x)


def quaternion_matrix3(q):
    """3x3 block of transformations.quaternion_matrix (w, x, y, z), as cost_functions.py:49 uses it"""
    q = np.array(q, dtype=np.float64, copy=True)
    n = float(np.dot(q, q))
    if n < _EPS4:
        return np.identity(3)
    q *= math.sqrt(2.0 / n)
    o = np.outer(q, q)
    return np.array([[1.0 - o[2, 2] - o[3, 3], o[1, 2] - o[3, 0], o[1, 3] + o[2, 0]],
                     [o[1, 2] + o[3, 0], 1.0 - o[1, 1] - o[3, 3], o[2, 3] - o[1, 0]],
                     [o[1, 3] - o[2, 0], o[2, 3] + o[1, 0], 1.0 - o[1, 1] - o[2, 2]]])


def init_gmm_params(X, k):
    """gmmreg_gpu/gmm_impl.py:18-24: KMeans(n_clusters=k, random_state=1, max_iter=50, n_init=1) centres, weights 1/k.
    scikit-learn is the reference's own dependency for this step (host side, once per fit)."""
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=k, random_state=1, max_iter=50, n_init=1).fit(np.asarray(X))
    return km.cluster_centers_, np.ones(k) / k


def train_gmm(X, max_iter, tol, means, covariances, weights=None, engine=None):
    """gmmreg_gpu/gmm_impl.py:62-83 (diagonal; clip(cov, 0); pi = nk/N) -> (inv_cov, means, weights, covariances, log_ll)"""
    eng = engine or _gi.default_engine()
    k = len(means)
    if weights is None:
        weights = np.ones(k) / k
    eng.set_points(np.asarray(X, dtype=np.float32))
    res = eng.fit_flat(_gi._host(means), _gi._host(covariances), _gi._host(weights), cov_type="diag", flavor=FLAVOR_PY_OLD,
                       max_iter=max_iter, tol=tol)
    lls = [float(v) for v in res["ll"]]
    if not (len(lls) >= 2 and abs(lls[-1] - lls[-2]) < tol):
        print('Failed to converge. Increase max-iter or tol.')
    return res["inv_cov"], res["means"], res["weights"], res["covs"], lls


class GMM_GPU_Base:
    """gmmreg_gpu/gmm.py:72-104"""

    def __init__(self, num_components, max_iter=30, tol=1e-4, engine=None, verbose=False):
        self.num_components = num_components
        self.max_iter = max_iter
        self.tol = tol
        self._engine = engine
        self._verbose = verbose

    def fit(self, X):
        X = _cv(X)
        means, weights = init_gmm_params(X, self.num_components)
        covs = 0.1 * np.ones((self.num_components, means.shape[1]), dtype=np.float32)
        out = train_gmm(X.astype(np.float32), self.max_iter, self.tol, means.astype(np.float32), covs, weights.astype(np.float32),
                        engine=self._engine)
        self.inv_covs, self.means_, self.weights_, self.covariances_, self.lls = out
        if self._verbose and len(self.lls):
            print("\nLog Likelihood Min-Max:\n\n", np.min(self.lls), np.max(self.lls))
        return self


class GMM_GPU:
    """gmmreg_gpu/gmm.py:44-57: compute(data) -> (means, weights) -- the covariances are dropped on this path"""

    def __init__(self, n_gmm_components=100, max_iter=30, tol=1e-4, engine=None):
        self._n_gmm_components = n_gmm_components
        self.max_iter = max_iter
        self.tol = tol
        self._engine = engine

    def init(self):
        self._clf = GMM_GPU_Base(self._n_gmm_components, max_iter=self.max_iter, tol=self.tol, engine=self._engine)

    def compute(self, data):
        self._clf.fit(data)
        return self._clf.means_, self._clf.weights_

    def annealing(self):
        pass

    def __call__(self, data):
        return self.compute(data)


class RigidCostFunction:
    """cost_functions.py:43-69; __call__(theta, mu_source, phi_source, mu_target, phi_target, sigma) -> (f, grad[7])
    evaluated by libhgmm (float64).  The mixtures are uploaded when they change (identity of the argument arrays)."""

    def __init__(self, engine=None):
        self._tf_type = RigidTransformation
        self._engine = engine
        self._key = None

    def to_transformation(self, theta):
        return self._tf_type(quaternion_matrix3(theta[:4]), np.array(theta[4:7], dtype=np.float64))

    def initial(self):
        x0 = np.zeros(7)
        x0[0] = 1.0
        return x0

    def _upload(self, mu_source, phi_source, mu_target, phi_target):
        eng = self._engine or _gi.default_engine()
        key = (id(mu_source), id(phi_source), id(mu_target), id(phi_target))
        if key != self._key:
            eng.l2_set_mixtures(mu_source, phi_source, mu_target, phi_target)
            self._key = key
            self._keep = (mu_source, phi_source, mu_target, phi_target)     # keeps the ids alive
        return eng

    def __call__(self, theta, *args):
        mu_source, phi_source, mu_target, phi_target, sigma = args
        eng = self._upload(mu_source, phi_source, mu_target, phi_target)
        return eng.l2_cost_grad(theta, sigma)

    def minimize_on_device(self, theta, args, max_iter, gtol):
        mu_source, phi_source, mu_target, phi_target, sigma = args
        eng = self._upload(mu_source, phi_source, mu_target, phi_target)
        return eng.l2_optimize(theta, sigma, max_iter=max_iter, gtol=gtol)


class L2DistRegistration(object):
    """gmmreg.py:15-121"""

    def __init__(self, source, feature_gen, cost_fn, sigma=1.0, delta=0.9, use_estimated_sigma=True, optimizer="scipy",
                 verbose=False):
        self._source = source
        self._feature_gen = feature_gen
        self._cost_fn = cost_fn
        self._sigma = sigma
        self._delta = delta
        self._use_estimated_sigma = use_estimated_sigma
        self._callbacks = []
        self._optimizer = optimizer
        self._verbose = verbose
        self.last_result = None
        if self._source is not None and self._use_estimated_sigma:
            self._estimate_sigma(self._source)

    def set_source(self, source):
        self._source = source
        if self._use_estimated_sigma:
            self._estimate_sigma(self._source)

    def set_callbacks(self, callbacks):
        self._callbacks.extend(callbacks)

    def _estimate_sigma(self, data):
        ndata, ndim = data.shape
        data_hat = data - np.mean(data, axis=0)
        self._sigma = np.power(np.linalg.det(np.dot(data_hat.T, data_hat) / (ndata - 1)), 1.0 / (2.0 * ndim))
        if self._verbose:
            print("Estimated Sigma: ", self._sigma)

    def _annealing(self):
        self._sigma *= self._delta

    def optimization_cb(self, x):
        tf_result = self._cost_fn.to_transformation(x)
        for c in self._callbacks:
            c(tf_result)

    def registration(self, target, maxiter=1, tol=1.0e-3, opt_maxiter=10, opt_tol=1.0e-5):
        start = time.time()
        f = None
        x_ini = self._cost_fn.initial()
        self._feature_gen.init()
        mu_target, phi_target = self._feature_gen.compute(target)
        mu_target = np.ascontiguousarray(mu_target, np.float64)
        phi_target = np.ascontiguousarray(phi_target, np.float64) * 1e3          # gmmreg.py:75
        x = x_ini
        fun = None
        for _ in range(maxiter):
            mu_source, phi_source = self._feature_gen.compute(self._source)
            mu_source = np.ascontiguousarray(mu_source, np.float64)
            phi_source = np.ascontiguousarray(phi_source, np.float64) * 1e3      # gmmreg.py:88
            args = (mu_source, phi_source, mu_target, phi_target, float(self._sigma))
            if self._optimizer == "scipy":
                from scipy.optimize import minimize
                res = minimize(self._cost_fn, x_ini, args=args, method='BFGS', jac=True, tol=opt_tol,
                               options={'maxiter': opt_maxiter}, callback=self.optimization_cb)
                x, fun = res.x, float(res.fun)
                self.last_result = {"x": x, "fun": fun, "nit": int(res.nit), "nfev": int(res.nfev), "status": int(res.status)}
            elif self._optimizer == "device":
                x, fun, nit, nfev, status = self._cost_fn.minimize_on_device(x_ini, args, opt_maxiter, opt_tol)
                self.optimization_cb(x)
                self.last_result = {"x": x, "fun": fun, "nit": nit, "nfev": nfev, "status": status}
            else:
                raise ValueError("optimizer must be 'scipy' or 'device'")
            self._annealing()
            self._feature_gen.annealing()
            if f is not None and abs(fun - f) < tol:
                break
            f = fun
            x_ini = x
        if self._verbose:
            print("Overall Time taken: ", time.time() - start)
        return self._cost_fn.to_transformation(x)


class RigidGMMReg(L2DistRegistration):
    """gmmreg.py:138-147"""

    def __init__(self, source, sigma=1.0, delta=0.9, n_gmm_components=50, use_estimated_sigma=True, optimizer="scipy",
                 engine=None, verbose=False):
        n_gmm_components = min(n_gmm_components, int(source.shape[0] * 0.8))
        super(RigidGMMReg, self).__init__(source, GMM_GPU(n_gmm_components, max_iter=10, engine=engine),
                                          RigidCostFunction(engine=engine), sigma, delta, use_estimated_sigma,
                                          optimizer=optimizer, verbose=verbose)


def registration_gmmreg(source, target, tf_type_name='rigid', callbacks=[], **kargs):
    """gmmreg.py:149-157"""
    if tf_type_name == 'rigid':
        gmmreg = RigidGMMReg(_cv(source), **kargs)
    else:
        raise ValueError('Unknown transform type %s' % tf_type_name)
    gmmreg.set_callbacks(callbacks)
    return gmmreg.registration(_cv(target))
