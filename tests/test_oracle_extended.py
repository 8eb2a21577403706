"""The extended-precision arbiter (oracle/extended.py) against the float64 oracle, on the CPU.

SURVEY 7.1 / 8(c): numpy.longdouble decides the 1e-12 entry tolerance per kernel kind and the 1e-9 tolerance on
energy / gradient at the stress noise.  Here: the float64 oracle's kernel entries are within 1e-12 (relative, per
entry) of the extended values for every kind on the path, its energy / gradient within 1e-9 at sigma^2 = 1e-6 --
i.e. the oracle the GPU tests compare with is itself inside the contract."""
import math

import numpy as np
import pytest

from oracle import extended as ext
from oracle import gp_oracle as orc

pytestmark = pytest.mark.skipif(not ext.available(), reason="numpy.longdouble is not wider than float64 here")

KINDS = {
    "se": orc.Kernel("se", {"bandwidth": 1.7, "amplitude": 1.2}),
    "rbf": orc.Kernel("rbf", {"bandwidth": 2.1}),
    "ard": orc.Kernel("ard", {"MahalanobisAmplitude": 1.1, **{f"MahalanobisBandwidth_{i}": 1.5 + 0.3 * i for i in range(6)}}),
    "matern1": orc.Kernel("matern", {"p": 1.0, "l": 2.2}),
    "matern2": orc.Kernel("matern", {"p": 2.0, "l": 2.2}),
    "matern3": orc.Kernel("matern", {"p": 3.0, "l": 2.2}),
    "periodic": orc.Kernel("periodic", {"sigma": 1.3, "lengthscale": 1.4, "frequency": 0.07}),
    "cauchy": orc.Kernel("cauchy", {"sigma": 2.3}),
    "laplacian": orc.Kernel("laplacian", {"beta": 3.1}),
    "ratquad": orc.Kernel("ratquad", {"mu": 1.5, "l": 1.9}),
}


@pytest.mark.parametrize("name", sorted(KINDS))
def test_float64_entries_are_within_1e12_of_extended(name):
    X, _, _ = orc.synthetic(160, 6, seed=11)
    cov = KINDS[name]
    K64 = cov.value(X, X)
    Kld = ext.kernel_matrix(cov, X)
    assert Kld.dtype == np.longdouble
    err = np.abs(K64 - Kld) / np.maximum(np.abs(Kld), 1e-290)
    assert float(err.max()) <= 1e-12
    # and the gradient matrices the trace pass re-creates
    G64, Gld = cov.grads(X, X), cov.grads(np.asarray(X, dtype=np.longdouble), np.asarray(X, dtype=np.longdouble))
    for k in G64:
        scale = float(np.abs(Gld[k]).max()) or 1.0
        assert float(np.abs(G64[k] - Gld[k]).max()) <= 1e-12 * scale, k


def test_extended_linear_algebra_is_consistent():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((60, 60))
    A = A @ A.T + 60 * np.eye(60)
    L = ext.cholesky(A)
    assert float(np.abs(L @ L.T - A).max()) <= 1e-17 * float(np.abs(A).max()) * 60
    b = rng.standard_normal(60)
    x = ext.solve_upper_t(L, ext.solve_lower(L, b))
    assert float(np.abs(A.astype(np.longdouble) @ x - b).max()) <= 1e-16
    with pytest.raises(np.linalg.LinAlgError):
        ext.cholesky(-np.eye(3))


@pytest.mark.parametrize("name,noise", [("se", 0.1), ("se", 1e-3), ("ard", 1e-3), ("matern2", 1e-3), ("periodic", 2.0)])
def test_float64_energy_and_gradient_within_1e9_of_extended(name, noise):
    """sigma = 1e-3 (sigma^2 = 1e-6) is the stress point the tolerance is quoted at."""
    X, y, _ = orc.synthetic(220, 6, seed=5)
    if name == "periodic":
        X = X[:, :3]   # the L1-norm periodic kernel is indefinite for d > 1 (min eigenvalue -1.3 here): large noise
    cov = KINDS[name]
    e, g = ext.energy_grad(cov, noise, X, y)
    eo, go = orc.energy(cov, noise, X, y), orc.grad_energy(cov, noise, X, y)
    assert abs(eo - float(e)) <= 1e-9 * abs(float(e))
    assert set(g) == set(go)
    for k, v in go.items():
        assert abs(v - float(g[k])) <= 1e-9 * abs(float(g[k])), (k, v, float(g[k]))


def test_extended_predictive_matches_oracle():
    X, y, Xs = orc.synthetic(150, 4, seed=2, m=12)
    cov = KINDS["matern2"]
    mu, cv = ext.predictive(cov, 0.2, X, y, Xs)
    mo, co = orc.predictive(cov, 0.2, X, y, Xs)
    assert float(np.abs(mo - mu).max()) <= 1e-12 * float(np.abs(mu).max())
    assert float(np.abs(co - cv).max()) <= 1e-12 * float(np.abs(cv).max())
