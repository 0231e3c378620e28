"""Tensor-core formulation of the flat E/M iteration, EMULATED on the CPU -- the experiment behind DESIGN.md 3.1's decision
(VERDICT r1 item 8).  Not a bench line; writes profiles/r02_tensor_core_experiment.json.

The north star asks for tensor cores "where the N x J point . Sigma^-1 . point contraction is large enough to tile as a dense
GEMM".  The GEMM form of configs[1] (bun000, J = 800, full covariance, 10 EM iterations, the bench's init) is
    E:  Q [N x J]  = Phi [N x 10] . W [10 x J]      Phi(x) = (1, x, y, z, xx, yy, zz, xy, xz, yz) of x' = x - c,  K = 10
    M:  S [J x 10] = Gamma^T [J x N] . Phi [N x 10]                                                               K = N
with the expansion point c either the cloud's centroid ("global") or the centroid of each CTA-sized chunk of 272 consecutive
points ("chunk": W's four affine rows are then rebuilt per chunk, the raw moments un-shifted per chunk in float64 -- what a
kernel would do).  Operands are rounded the way a tensor-core pipeline sees them, products accumulate in float32:
    fp32     both operands float32 (the formulation's own cancellation, no tensor core)
    tf32     one TF32 term (10-bit mantissa)          tf32x3  a = a1 + a2, three cross products (the usual "3xTF32")
    bf16     one BF16 term (7-bit mantissa)           bf16x3  a = a1 + a2 + a3, six cross products ("BF16x3", fp32-class)
Each variant runs the SAME 10 iterations from the same start; the table is the relative Frobenius distance of pi / mu / Sigma
to the float64 oracle (oracle/flat_gmm.py::cpp_fit), next to the shipped kernel's own figure for the same fit.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flat_gmm      # noqa: E402

CHUNK = 272


def rnd(a, keep_bits):
    """round-to-nearest-even of float32 `a` to `keep_bits` explicit mantissa bits (TF32: 10, BF16: 7)"""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    drop = 23 - keep_bits
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = (u + half - np.uint64(1) + lsb) >> np.uint64(drop) << np.uint64(drop)
    return u.astype(np.uint32).view(np.float32)


def split(a, bits, terms):
    out, r = [], np.asarray(a, np.float32)
    for _ in range(terms):
        h = rnd(r, bits)
        out.append(h)
        r = (r - h).astype(np.float32)
    return out


def mm(A, B, mode):
    """A [m,k] @ B [k,n] with tensor-core operand rounding, float32 accumulation"""
    A = np.asarray(A, np.float32)
    B = np.asarray(B, np.float32)
    if mode == "fp32":
        return A @ B
    bits = 10 if mode.startswith("tf32") else 7
    if mode in ("tf32", "bf16"):
        return rnd(A, bits) @ rnd(B, bits)
    if mode == "tf32x3":
        a, b = split(A, bits, 2), split(B, bits, 2)
        return (a[1] @ b[0] + a[0] @ b[1]) + a[0] @ b[0]
    if mode == "bf16x3":
        a, b = split(A, bits, 3), split(B, bits, 3)
        return (((a[2] @ b[0] + a[1] @ b[1]) + a[0] @ b[2]) + (a[1] @ b[0] + a[0] @ b[1])) + a[0] @ b[0]
    raise ValueError(mode)


def features(Xc):
    x, y, z = Xc[:, 0], Xc[:, 1], Xc[:, 2]
    return np.stack([np.ones_like(x), x, y, z, x * x, y * y, z * z, x * y, x * z, y * z], 1).astype(np.float32)


def weights(logpi, mu, cov, c):
    """W [10, J] of q_j(x') = log(pi_j N_j(x' + c)) as a polynomial in x' (float64 -> float32, as a finalize kernel would)"""
    P = np.linalg.inv(cov)
    m = mu - c[None, :]
    const = logpi - 0.5 * (3 * np.log(2 * np.pi) + np.log(np.linalg.det(cov))) - 0.5 * np.einsum("ja,jab,jb->j", m, P, m)
    lin = np.einsum("jab,jb->ja", P, m)
    W = np.stack([const, lin[:, 0], lin[:, 1], lin[:, 2], -0.5 * P[:, 0, 0], -0.5 * P[:, 1, 1], -0.5 * P[:, 2, 2],
                  -P[:, 0, 1], -P[:, 0, 2], -P[:, 1, 2]], 0)
    return W.astype(np.float32)


def em_gemm(X, mu0, iters, s0, mode, centre):
    N, J = len(X), len(mu0)
    mu = np.array(mu0, np.float64)
    cov = np.tile(np.eye(3) * s0, (J, 1, 1))
    logpi = np.full(J, -np.log(J))
    X64 = X.astype(np.float64)
    chunks = [(0, N)] if centre == "global" else [(s, min(N, s + CHUNK)) for s in range(0, N, CHUNK)]
    for _ in range(iters):
        S0 = np.zeros(J)
        S1 = np.zeros((J, 3))
        S2 = np.zeros((J, 3, 3))
        for (a, b) in chunks:
            c = X64[a:b].mean(axis=0).astype(np.float32).astype(np.float64)
            Phi = features((X[a:b] - c.astype(np.float32)).astype(np.float32))
            Q = mm(Phi, weights(logpi, mu, cov, c), mode)                       # E: K = 10
            Q = Q - Q.max(axis=1, keepdims=True)
            G = np.exp(Q)
            G = (G / G.sum(axis=1, keepdims=True)).astype(np.float32)
            S = mm(G.T, Phi, mode).astype(np.float64)                            # M: K = N (this chunk)
            s0_, s1_ = S[:, 0], S[:, 1:4]
            s2_ = np.empty((J, 3, 3))
            s2_[:, 0, 0], s2_[:, 1, 1], s2_[:, 2, 2] = S[:, 4], S[:, 5], S[:, 6]
            s2_[:, 0, 1] = s2_[:, 1, 0] = S[:, 7]
            s2_[:, 0, 2] = s2_[:, 2, 0] = S[:, 8]
            s2_[:, 1, 2] = s2_[:, 2, 1] = S[:, 9]
            # un-shift the chunk's raw moments about c to moments about the origin, float64
            S0 += s0_
            S1 += s1_ + s0_[:, None] * c[None, :]
            S2 += (s2_ + s1_[:, :, None] * c[None, None, :] + c[None, :, None] * s1_[:, None, :]
                   + s0_[:, None, None] * (c[:, None] * c[None, :])[None])
        pi = S0 / S0.sum()
        mu = S1 / S0[:, None]
        cov = S2 / S0[:, None, None] - mu[:, :, None] * mu[:, None, :]
        logpi = np.log(pi)
    return pi, mu, cov


def main():
    X = np.load(os.path.join(ROOT, "tests", "golden", "bun000_xyz.npy")).astype(np.float32)
    J, iters, s0 = 800, 10, 1e-4
    mu0 = X[np.random.default_rng(1).choice(len(X), J, replace=False)]
    t0 = time.time()
    ow, omu, ocov, _ = flat_gmm.cpp_fit(X, mu0, iters, sigma0_sq=s0)
    print("oracle: %.0f s" % (time.time() - t0), flush=True)
    rows = []
    for centre in ("global", "chunk"):
        for mode in ("fp32", "tf32", "tf32x3", "bf16", "bf16x3"):
            t0 = time.time()
            try:
                pi, mu, cov = em_gemm(X, mu0, iters, s0, mode, centre)
                e = {"pi": flat_gmm.rel_fro(pi, ow), "mu": flat_gmm.rel_fro(mu, omu), "cov": flat_gmm.rel_fro(cov, ocov)}
                ok = bool(np.isfinite(list(e.values())).all() and max(e.values()) < 1e-4)
            except Exception as ex:      # singular covariances after a diverged iteration
                e, ok = {"error": repr(ex)}, False
            rows.append({"expansion_point": centre, "operands": mode, "rel_fro_vs_float64_oracle": e, "meets_1e-4": ok,
                         "seconds": time.time() - t0})
            print(rows[-1], flush=True)
    out = {"workload": "configs[1]: bun000 (40256 pts), J=800 full covariance, 10 EM iterations, init seed 1, Sigma0 = 1e-4 I",
           "emulation": "NumPy: operands rounded to TF32 / BF16 terms (round-to-nearest-even), products accumulated in float32",
           "shipped_kernel_same_fit": "em_flat7 (FFMA2, direct d^T A d about the component mean): 2e-7 .. 7e-6 (tests/test_gpu_parity.py)",
           "rows": rows}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_tensor_core_experiment.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
