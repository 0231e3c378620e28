"""CPU: the algebra em_flat8_kernel (csrc/flat_em8.cu) relies on -- sums taken about a chunk origin o and moved to a component
mean m equal the sums taken about m directly (the reference's maximizationStep accumulates about the data origin and subtracts
mu mu^T, gmm_kernels.cu:135-210; both are the same second central moment)."""
import numpy as np


def test_recentring_identity_float64():
    rng = np.random.default_rng(0)
    x = rng.normal(0.0, 0.02, (32, 3)) + np.array([0.1, -0.05, 0.07])
    g = rng.random(32)
    m = np.array([0.11, -0.04, 0.06])
    o = x[16]
    u = x - o
    S0, S1 = g.sum(), (g[:, None] * u).sum(0)
    S2 = np.einsum("p,pa,pb->ab", g, u, u)
    nd = o - m                                             # the kernel's nd = -delta
    M1 = S1 + nd * S0
    M2 = np.empty((3, 3))
    for a in range(3):
        for b in range(a, 3):
            M2[a, b] = M2[b, a] = S2[a, b] + nd[a] * S1[b] + nd[b] * M1[a]      # two FMAs per entry, as in the kernel
    d = x - m
    assert np.allclose(M1, (g[:, None] * d).sum(0), rtol=0, atol=1e-15)
    assert np.allclose(M2, np.einsum("p,pa,pb->ab", g, d, d), rtol=0, atol=1e-16)


def test_cholesky_form_equals_quadratic_form():
    """e = 2^-(|L^T u + b|^2 + k4) with -A = L L^T, b = L^T (o - m), k4 = Cref - c2 is 2^(c2 + d^T A d - Cref)"""
    rng = np.random.default_rng(1)
    B = rng.normal(size=(3, 3))
    Minus_A = B @ B.T + 0.1 * np.eye(3)                    # -A, positive definite
    L = np.linalg.cholesky(Minus_A)
    m, o = rng.normal(size=3), rng.normal(size=3)
    c2, cref = -3.0, 1.5
    for _ in range(10):
        x = rng.normal(size=3)
        d = x - m
        q = c2 - d @ Minus_A @ d - cref
        r = L.T @ (x - o) + L.T @ (o - m)
        assert abs(-(r @ r + (cref - c2)) - q) < 1e-12
