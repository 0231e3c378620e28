"""Feature-style classes of the reference (src/python/gmm_waymo/src/gmm.py, gmmreg_gpu/gmm.py).

Same class names, constructor arguments, methods and attributes; the fit runs in libhgmm.
`GMM_CPU*` are kept as aliases of the GPU classes (the reference's CPU classes run the same
`train_gmm` on NumPy arrays, gmm.py:120-146) -- there is deliberately no CPU code path here.
"""
import abc

import numpy as np

from .gmm_impl import train_gmm, init_gmm_params, timer, predict


class Feature(abc.ABC):
    """gmm.py:14-27"""

    @abc.abstractmethod
    def init(self):
        pass

    @abc.abstractmethod
    def compute(self, data):
        return None

    def annealing(self):
        pass

    def __call__(self, data):
        return self.compute(data)


class GMM_GPU_Base:
    """gmm.py:65-101: fit() draws the reference's random init, trains, and exposes
    means_, covariances_, weights_, lls, inv_covs."""

    def __init__(self, num_components, max_iter=30, tol=1e-4, cov_type='diag', engine=None, rng=None, verbose=True):
        self.num_components = num_components
        self.max_iter = max_iter
        self.tol = tol
        self.cov_type = cov_type
        self._engine = engine
        self._rng = rng
        self._verbose = verbose

    def fit(self, X):
        X = np.asarray(X.points if hasattr(X, "points") else X).astype(np.float32)
        means, weights, covs = init_gmm_params(X, self.num_components, cov_type=self.cov_type, rng=self._rng)
        if self._verbose:
            with timer('GPU GMM TRAIN'):
                out = train_gmm(X, self.max_iter, self.tol, means, covs, weights, cov_type=self.cov_type, engine=self._engine)
        else:
            out = train_gmm(X, self.max_iter, self.tol, means, covs, weights, cov_type=self.cov_type, engine=self._engine)
        self.inv_covs, self.means_, self.weights_, self.covariances_, self.lls = out
        if self._verbose and len(self.lls):
            print("\nLog Likelihood Min-Max:\n\n", np.min(self.lls), np.max(self.lls))
        return self

    def predict(self, X):
        X = np.asarray(X.points if hasattr(X, "points") else X).astype(np.float32)
        return predict(X, self.inv_covs, self.means_, self.weights_, cov_type=self.cov_type, engine=self._engine)


class GMM_GPU(Feature):
    """gmm.py:46-63: compute(data) -> (means, weights, covariances, inv_covs)."""

    def __init__(self, n_gmm_components=100, max_iter=30, tol=1e-4, cov_type='diag', engine=None, rng=None, verbose=True):
        self._n_gmm_components = n_gmm_components
        self.max_iter = max_iter
        self.tol = tol
        self.cov_type = cov_type
        self._engine = engine
        self._rng = rng
        self._verbose = verbose

    def init(self):
        self._clf = GMM_GPU_Base(self._n_gmm_components, max_iter=self.max_iter, tol=self.tol, cov_type=self.cov_type,
                                 engine=self._engine, rng=self._rng, verbose=self._verbose)

    def compute(self, data):
        self._clf.fit(data)
        return self._clf.means_, self._clf.weights_, self._clf.covariances_, self._clf.inv_covs

    def predict(self, data):
        return self._clf.predict(data)


class GMM_CPU_Base(GMM_GPU_Base):
    """gmm.py:120-146 (same train_gmm; here it is the same device engine)."""


class GMM_CPU(GMM_GPU):
    """gmm.py:103-118: compute(data) -> (means, weights, covariances)."""

    def compute(self, data):
        self._clf.fit(data)
        return self._clf.means_, self._clf.weights_, self._clf.covariances_
