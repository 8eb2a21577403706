"""Extended-precision arbiter for the GP hot path -- TEST INFRASTRUCTURE ONLY (see oracle/gp_oracle.py).

SURVEY.md 7.1 / 8(c): the float64 oracle and the device are two independent float64 evaluations of the same
formulas; where they differ by more than the stated tolerance (the gradient at the stress noise sigma^2 = 1e-6,
where cond(K) ~ 1e6 amplifies rounding to ~1e-8), a third evaluation in higher precision decides which side is
off and by how much.  Everything here runs in numpy.longdouble (x87 80-bit extended on this platform: 64-bit
mantissa, eps = 1.08e-19, i.e. 2048 times finer than float64) with no LAPACK underneath: kernel entries from the
oracle's own elementwise code evaluated on longdouble inputs (direct differences, exp / sqrt in extended
precision), an unblocked right-looking Cholesky, triangular substitution, K^-1 = L^-T L^-1, and the reference's
energy / gradient formulas (AbstractGPRegressionModel.scala:231-241, 527-529).  O(n^3) NumPy loops: n <= ~600.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

from . import gp_oracle as orc

LD = np.longdouble


def available() -> bool:
    """True when longdouble really is wider than float64 (x86: 80-bit extended)."""
    return np.finfo(LD).eps < 1e-18


def kernel_matrix(cov, X: np.ndarray, noise_level: float | None = None) -> np.ndarray:
    """k(X, X) [+ Dirac noise] with every elementwise operation in longdouble."""
    Xl = np.asarray(X, dtype=LD)
    K = cov.value(Xl, Xl)
    assert K.dtype == LD
    if noise_level is not None:
        K = K + orc.dirac(noise_level).value(Xl, Xl)
    return K


def cross_matrix(cov, X: np.ndarray, Y: np.ndarray) -> np.ndarray:
    return cov.value(np.asarray(X, dtype=LD), np.asarray(Y, dtype=LD))


def cholesky(A: np.ndarray) -> np.ndarray:
    """Unblocked right-looking Cholesky in longdouble (lower factor)."""
    A = np.array(A, dtype=LD)
    n = A.shape[0]
    for k in range(n):
        d = A[k, k]
        if not d > 0:
            raise np.linalg.LinAlgError(f"pivot {k} is not positive")
        lkk = np.sqrt(d)
        A[k, k] = lkk
        if k + 1 < n:
            col = A[k + 1:, k] / lkk
            A[k + 1:, k] = col
            A[k + 1:, k + 1:] -= np.outer(col, col)
    return np.tril(A)


def solve_lower(L: np.ndarray, B: np.ndarray) -> np.ndarray:
    """L^-1 B by forward substitution (B: vector or matrix)."""
    X = np.array(B, dtype=LD)
    n = L.shape[0]
    for k in range(n):
        X[k] = X[k] / L[k, k]
        if k + 1 < n:
            X[k + 1:] -= np.multiply.outer(L[k + 1:, k], X[k]) if X.ndim == 2 else L[k + 1:, k] * X[k]
    return X


def solve_upper_t(L: np.ndarray, B: np.ndarray) -> np.ndarray:
    """L^-T B by back substitution."""
    X = np.array(B, dtype=LD)
    n = L.shape[0]
    for k in range(n - 1, -1, -1):
        X[k] = X[k] / L[k, k]
        if k > 0:
            X[:k] -= np.multiply.outer(L[k, :k], X[k]) if X.ndim == 2 else L[k, :k] * X[k]
    return X


def energy_grad(cov, noise_level: float, X: np.ndarray, y: np.ndarray, mean: np.ndarray | None = None
                ) -> Tuple[LD, Dict[str, LD]]:
    """(E, g) in extended precision: E = 0.5 (r' K^-1 r + sum log L_ii + n log 2 pi);
    g_theta = sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij / dtheta for every effective theta, then the noise level."""
    n = X.shape[0]
    Xl = np.asarray(X, dtype=LD)
    r = np.asarray(y, dtype=LD) - (0 if mean is None else np.asarray(mean, dtype=LD))
    nz = orc.dirac(noise_level)
    K = cov.value(Xl, Xl) + nz.value(Xl, Xl)
    L = cholesky(K)
    z = solve_lower(L, r)
    alpha = solve_upper_t(L, z)
    two_pi = 2 * LD(math.pi) if not hasattr(np, "pi") else 2 * np.arccos(LD(-1))
    e = (z @ z + np.log(np.diag(L)).sum() + n * np.log(two_pi)) / 2
    W = solve_lower(L, np.eye(n, dtype=LD))     # L^-1
    kinv = W.T @ W
    M = np.outer(alpha, alpha) - kinv
    g = {name: (M * dK).sum() for name, dK in cov.grads(Xl, Xl).items() if name in cov.effective()}
    g["noiseLevel"] = (M * nz.grads(Xl, Xl)["noiseLevel"]).sum()
    return e, g


def predictive(cov, noise_level: float, X: np.ndarray, y: np.ndarray, Xs: np.ndarray):
    """Posterior mean and covariance (AbstractGPRegressionModel.scala:434-466) in extended precision."""
    Xl, Xsl = np.asarray(X, dtype=LD), np.asarray(Xs, dtype=LD)
    K = cov.value(Xl, Xl) + orc.dirac(noise_level).value(Xl, Xl)
    L = cholesky(K)
    alpha = solve_upper_t(L, solve_lower(L, np.asarray(y, dtype=LD)))
    Ks = cov.value(Xl, Xsl)
    V = solve_lower(L, Ks)
    return Ks.T @ alpha, cov.value(Xsl, Xsl) - V.T @ V


def error_bars(mu: np.ndarray, cov_post: np.ndarray, sigma: float):
    """mean -/+ sigma * (chol(Sigma*) . 1): the reference's row-sum bars (BlockedMultiVariateGaussian.scala:58-68)."""
    bar = cholesky(cov_post) @ np.ones(cov_post.shape[0], dtype=LD)
    return mu - sigma * bar, mu + sigma * bar
