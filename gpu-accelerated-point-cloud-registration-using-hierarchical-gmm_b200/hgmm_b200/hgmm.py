"""Hierarchical-GMM tree build + tree registration with the reference's surface
(src/python/hgmm/hgmm_gpu.py; degenerate-case semantics of hgmm_cupy_cpu_working.py).

    buildGMMTree(points, maxTreeLevel, ls, ld) -> (mixingCoeff[nTotal], mean[nTotal,3], covar[nTotal,3,3])
    GMMTree(source, tree_level=5, lambda_c=0.01).registration(target, maxiter=20, tol=1e-4)
    registration_gmmtree(source, target, maxiter=20, tol=1e-4, callbacks=[], **kargs)

Everything numeric happens in libhgmm; this file only reproduces the reference's host-side
conventions (seeded init draw, the stateful transform of GMMTree, the returned inverse transform).
"""
import abc
from collections import namedtuple

import numpy as np

from .engine import Engine
from .gmm_impl import default_engine

eps = 1.0e-15      # hgmm_gpu.py:29
n_node = 8         # hgmm_gpu.py:30


def child(j):
    """hgmm_gpu.py:84-89"""
    return (j + 1) * n_node


def level(l):
    """hgmm_gpu.py:91-92"""
    return n_node * (n_node ** l - 1) // (n_node - 1)


def reference_init_indices(maxTreeLevel, seed=72):
    """hgmm_gpu.py:467-470: np.random.seed(72); idxs = np.random.randint(nTotal, size=nTotal)
    (legacy MT19937 stream, reproduced with a private RandomState so the global RNG is untouched)."""
    nTotal = level(maxTreeLevel)
    return np.random.RandomState(seed).randint(nTotal, size=nTotal)


def deepest_live_node(current, mixingCoeff, maxTreeLevel):
    """node that actually holds each point in a pruned (ragged) tree: `current` is the leaf-level id the build reports; under a
    terminal node the children are blank (pi = 0), so walk up (parent(j) = j // 8 - 1, hgmm_gpu.py:84-89 inverted) to the
    first node that carries mass."""
    cur = np.asarray(current, dtype=np.int64).copy()
    pi = np.asarray(mixingCoeff)
    for _ in range(maxTreeLevel - 1):
        up = (pi[cur] <= 0) & (cur >= n_node)
        cur[up] = cur[up] // n_node - 1
    return cur


def _cloud(x):
    return np.asarray(x.points if hasattr(x, "points") else x)


def buildGMMTree(points, maxTreeLevel, ls, ld, sig2=0.004, init_means=None, ll_mode="level", engine=None,
                 return_details=False, **kw):
    """hgmm_gpu.py:466-548.  Extra keyword arguments expose what the reference hard-codes:
    sig2 (0.004, :477), init_means (default: points[reference_init_indices]), ll_mode
    ('level' = the reference's whole-level log-likelihood; 'estep' = fast mode)."""
    pts = _cloud(points).astype(np.float32)
    nTotal = level(maxTreeLevel)
    if init_means is None:
        if pts.shape[0] < nTotal:
            raise ValueError("buildGMMTree needs at least nTotal=%d points to draw its init (hgmm_gpu.py:470,489)" % nTotal)
        init_means = pts[reference_init_indices(maxTreeLevel)]
    eng = engine or default_engine()
    eng.set_points(pts)
    res = eng.fit_tree(init_means, maxTreeLevel, ls=ls, ld=ld, sig2=sig2, ll_mode=ll_mode, want_current=return_details, **kw)
    eng._tree_level = maxTreeLevel
    if return_details:
        return res["pi"], res["mu"], res["cov"], res
    return res["pi"], res["mu"], res["cov"]


class Transformation(abc.ABC):
    """hgmm_gpu.py:583-597"""

    def transform(self, points, array_type=None):
        if array_type is not None and isinstance(points, array_type):
            return array_type(self._transform(np.asarray(points)))
        return self._transform(points)

    @abc.abstractmethod
    def _transform(self, points):
        return points


class RigidTransformation(Transformation):
    """hgmm_gpu.py:599-618"""

    def __init__(self, rot=np.identity(3), t=np.zeros(3), scale=1.0):
        self.rot = rot
        self.t = t
        self.scale = scale

    def _transform(self, points):
        return self.scale * np.dot(points, self.rot.T) + self.t

    def inverse(self):
        return RigidTransformation(self.rot.T, -np.dot(self.rot.T, self.t), 1.0 / self.scale)


EstepResult = namedtuple('EstepResult', ['momentZero', 'momentOne', 'momentTwo'])     # hgmm_gpu.py:666
MstepResult = namedtuple('MstepResult', ['transformation', 'q'])                      # hgmm_gpu.py:667


class GMMTree():
    """hgmm_gpu.py:669-768.  The constructor builds the tree on `source` once (the reference builds it
    three times for timing, :688-705); `_tf_result` persists across calls like the reference's."""

    def __init__(self, source=None, tree_level=5, lambda_c=0.01, ls=20, ld=1.0e-4, sig2=0.004, ll_mode="level",
                 solver="twist_lstsq", engine=None):
        self._source = None
        self._tree_level = tree_level
        self._lambda_c = lambda_c
        self._ls, self._ld, self._sig2, self._ll_mode = ls, ld, sig2, ll_mode
        self._solver = solver
        self._tf_type = RigidTransformation
        self._tf_result = self._tf_type()
        self._callbacks = []
        self._engine = engine or default_engine()
        self._target = None           # the array last uploaded as target (a reference is kept: `is` cannot alias a freed object)
        if source is not None:
            self.set_source(source)

    # The reference keeps the model per GMMTree object (hgmm_gpu.py:685-706).  Here the device copy lives in an Engine that other
    # GMMTree / buildGMMTree / predict calls may share: the host copy of the model is the object's own, and `model_token` /
    # `target_token` on the engine say who installed the device copy last -- anyone else re-installs before using it.
    def set_source(self, source):
        self._source = _cloud(source)
        self._mixingCoeff, self._mean, self._covar = buildGMMTree(self._source, self._tree_level, self._ls, self._ld,
                                                                  sig2=self._sig2, ll_mode=self._ll_mode, engine=self._engine)
        self._engine.model_token = self

    def set_model(self, mixingCoeff, mean, covar):
        """install an existing tree instead of building one"""
        self._mixingCoeff, self._mean, self._covar = (np.asarray(mixingCoeff), np.asarray(mean), np.asarray(covar))
        self._engine.tree_set_model(self._tree_level, self._mixingCoeff, self._mean, self._covar)
        self._engine.model_token = self

    def set_callbacks(self, callbacks):
        self._callbacks = callbacks

    def _ensure_model(self):
        if self._engine.model_token is not self:
            self._engine.tree_set_model(self._tree_level, self._mixingCoeff, self._mean, self._covar)
            self._engine.model_token = self

    def _ensure_target(self, target):
        """upload unless this very object uploaded this very array last (pass a copy, or call `invalidate_target()`, after
        mutating an array in place)"""
        if self._engine.target_token is not self or self._target is not target:
            self._engine.reg_set_target(target)
            self._engine.target_token = self
            self._target = target

    def invalidate_target(self):
        self._target = None

    def expectation_step(self, target):
        """hgmm_gpu.py:722-727: `target` is the ALREADY transformed cloud."""
        self._ensure_model()
        self._engine.reg_set_target(_cloud(target))
        self._target = None
        m0, m1, m2 = self._engine.reg_estep(np.identity(3), np.zeros(3), self._lambda_c, len(self._mixingCoeff))
        return EstepResult(m0, m1, m2)

    def maximization_step(self, estep_res, trans_p):
        """hgmm_gpu.py:729-752 on the moments of the last expectation_step."""
        rot, t, q = self._engine.reg_mstep(trans_p.rot, trans_p.t, solver=self._solver)
        return MstepResult(RigidTransformation(rot, t), np.array([q]))

    def registration(self, target, maxiter=20, tol=1.0e-4):
        """hgmm_gpu.py:754-768: iterate on the device from the current `_tf_result`; returns the
        inverse transform and the last q, like the reference."""
        tgt = _cloud(target)
        self._ensure_model()
        self._ensure_target(tgt)
        if self._callbacks:
            # per-iteration callbacks need the host in the loop
            q = None
            res = None
            for _ in range(maxiter):
                rot, t, qq, _, _ = self._engine.register_tree(self._tf_result.rot, self._tf_result.t, solver=self._solver,
                                                              maxiter=1, tol=0.0, lambda_c=self._lambda_c)
                res = MstepResult(RigidTransformation(rot, t), np.array([qq]))
                self._tf_result = res.transformation
                for c in self._callbacks:
                    c(self._tf_result.inverse())
                if q is not None and abs(qq - q) < tol:
                    break
                q = qq
            return MstepResult(self._tf_result.inverse(), res.q)
        rot, t, q, iters, _ = self._engine.register_tree(self._tf_result.rot, self._tf_result.t, solver=self._solver,
                                                         maxiter=maxiter, tol=tol, lambda_c=self._lambda_c)
        self._tf_result = RigidTransformation(rot, t)
        self.n_iter_ = iters
        return MstepResult(self._tf_result.inverse(), np.array([q]))


def registration_gmmtree(source, target, maxiter=20, tol=1.0e-4, callbacks=[], **kargs):
    """hgmm_gpu.py:802-807"""
    gt = GMMTree(_cloud(source), **kargs)
    gt.set_callbacks(callbacks)
    return gt.registration(_cloud(target), maxiter, tol)
