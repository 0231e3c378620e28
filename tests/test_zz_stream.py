"""GPU: the streaming segmentation loop (run_gmm_waymo_gpu.py:39-50 -- re-fit every k frames, hard-assign every frame)
against the oracle's train_gmm / predict restatements."""
import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu


def _frames(n_frames, n=6000, seed=0):
    rng = np.random.default_rng(seed)
    centres = np.array([[0.0, 0.0, 0.0], [0.4, 0.1, 0.0], [0.1, 0.5, 0.2], [-0.3, 0.2, 0.1]])
    which = rng.integers(0, len(centres), n)
    base = centres[which] + rng.normal(0, 0.03, (n, 3))
    return [(base + 0.004 * k * np.array([1.0, -0.5, 0.25]) + rng.normal(0, 1e-3, (n, 3))).astype(np.float32) for k in range(n_frames)], centres


@pytest.mark.parametrize("cov_type", ["diag", "spherical"])
def test_streaming_segmenter_matches_oracle(engine, cov_type):
    from hgmm_b200.stream import StreamingSegmenter
    from oracle import flat_gmm
    frames, centres = _frames(12)
    J = 4
    rng = np.random.default_rng(3)
    means0 = (centres + rng.normal(0, 0.05, (J, 3))).astype(np.float32)
    covs0 = np.full((J, 3) if cov_type == "diag" else (J,), 2e-3, np.float32)
    w0 = np.full(J, 1.0 / J, np.float32)
    seg = StreamingSegmenter(n_components=J, max_iter=60, tol=1e-4, cov_type=cov_type, fit_every=5, warm_start=True, engine=engine)
    seg.model = (None, means0, w0, covs0)               # start of the stream: a caller-supplied model instead of the random draw
    model = (means0, covs0, w0)
    for k, f in enumerate(frames):
        labels = seg.step(f)
        assert labels.shape == (len(f),) and labels.dtype == np.int32
        if k % 5 == 0:                                  # a re-fit frame: same EM as the oracle from the same start
            o = flat_gmm.py_train_gmm(f, 60, 1e-4, model[0], model[1], model[2], cov_type)
            assert seg.fit_iterations[-1] == len(o[4])
            assert rel_fro(seg.model[1], o[1]) < 1e-4 and rel_fro(seg.model[3], o[3]) < 1e-4 and rel_fro(seg.model[2], o[2]) < 1e-4
            model = (seg.model[1], seg.model[3], seg.model[2])
        inv, mu, w, _ = seg.model
        want = flat_gmm.py_predict(f.astype(np.float64), inv.astype(np.float64), mu.astype(np.float64), w.astype(np.float64), cov_type)
        assert (labels == want).mean() > 0.999, (k, (labels == want).mean())
    assert len(seg.fit_iterations) == 3
    assert seg.fit_iterations[1] <= seg.fit_iterations[0] and seg.fit_iterations[2] <= seg.fit_iterations[0]   # warm start pays


def test_streaming_segmenter_reference_defaults_run(engine):
    """the reference's own configuration (50 spherical components, random-scalar init, cold re-fit every 10 frames)"""
    from hgmm_b200.stream import segment_stream
    frames, _ = _frames(11, n=4000, seed=5)
    labels, seg = segment_stream(frames, n_components=50, max_iter=50, cov_type="spherical", fit_every=10, engine=engine,
                                 rng=np.random.default_rng(5))
    assert len(labels) == 11 and len(seg.fit_iterations) == 2
    assert all(l.min() >= 0 and l.max() < 50 for l in labels)
