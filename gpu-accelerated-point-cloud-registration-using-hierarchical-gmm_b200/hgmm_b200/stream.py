"""Frame-by-frame segmentation with periodic re-fit -- the loop of the reference's streaming driver
(src/python/gmm_waymo/src/run_gmm_waymo_gpu.py:39-50): every `fit_every` frames a flat mixture is fitted to the frame
(GMM_GPU.compute -> train_gmm), every frame is hard-assigned with the latest model (predict, gmm_impl.py:147-155:
argmax_j(log N_j(x) + log(pi_j + eps))).

    seg = StreamingSegmenter(n_components=50, max_iter=50, cov_type='spherical', fit_every=10)
    for frame in frames:                       # [N,3] float32 (NumPy, or a torch CUDA tensor: no host round trip)
        labels = seg.step(frame)               # int32 [N]

Differences from the reference loop, both opt-in: `warm_start=True` starts a re-fit from the previous model instead of a
fresh random draw (the scene changes little between re-fits, so EM stops on `tol` after a few iterations), and labels of a
frame that is not being fitted are computed against the resident model without re-installing it (Engine.predict_flat).
All compute runs in libhgmm (fused E+M sweep, hard-assignment kernel); nothing here has a CPU path.
"""
import numpy as np

from . import gmm_impl


class StreamingSegmenter:
    def __init__(self, n_components=50, max_iter=50, tol=1e-4, cov_type="spherical", fit_every=10, warm_start=False,
                 engine=None, rng=None):
        if cov_type not in ("diag", "spherical"):
            raise ValueError("cov_type must be 'diag' or 'spherical' (gmm_impl.py)")
        if fit_every < 1:
            raise ValueError("fit_every must be >= 1")
        self.n_components = int(n_components)
        self.max_iter = int(max_iter)
        self.tol = float(tol)
        self.cov_type = cov_type
        self.fit_every = int(fit_every)
        self.warm_start = bool(warm_start)
        self._engine = engine
        self._rng = rng
        self.frame_index = 0
        self.model = None            # (inv_cov, means, weights, covariances)
        self.lls = []                # mean log-likelihood per EM iteration of the latest fit
        self.fit_iterations = []     # EM iterations of every fit so far

    def _eng(self):
        return self._engine or gmm_impl.default_engine()

    def fit(self, frame):
        """(re)fit on `frame`; returns the labels of that frame."""
        eng = self._eng()
        if self.warm_start and self.model is not None:
            _, means, weights, covs = self.model
        else:
            means, weights, covs = gmm_impl.init_gmm_params(frame, self.n_components, cov_type=self.cov_type, rng=self._rng)
        eng.set_points(frame)
        res = eng.fit_flat(gmm_impl._host(means), gmm_impl._host(covs), gmm_impl._host(weights), cov_type=self.cov_type,
                           max_iter=self.max_iter, tol=self.tol)
        self.model = (res["inv_cov"], res["means"], res["weights"], res["covs"])
        self.lls = [float(v) for v in res["ll"]]
        self.fit_iterations.append(int(res["iters"]))
        return eng.predict_flat()                       # the frame just fitted is resident

    def step(self, frame):
        """labels (int32 [N]) of the next frame of the stream; re-fits on every `fit_every`-th frame."""
        if self.frame_index % self.fit_every == 0 or self.model is None:
            labels = self.fit(frame)
        else:
            labels = self._eng().predict_flat(frame)    # resident model, this frame's points
        self.frame_index += 1
        return labels

    def run(self, frames):
        for f in frames:
            yield self.step(f)


def segment_stream(frames, **kw):
    """labels of every frame of an iterable of [N,3] clouds (StreamingSegmenter with the same keyword arguments)."""
    seg = StreamingSegmenter(**kw)
    return [lab for lab in seg.run(frames)], seg
