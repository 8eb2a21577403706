"""Multi-output GP regression over the product index set (MOGPRegressionModel.scala)."""
import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _problem(n, d, q, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    W = rng.standard_normal((d, q)) / np.sqrt(d)
    Y = np.sin(X @ W) + 0.1 * rng.standard_normal((n, q))
    return X, Y


class _CoRegRBFOracle:
    """exp(-(i-j)^2 / b) with the reference's derivative k (i-j)^2 / |b|^3 (RBFKernel.scala:182-194)."""

    def __init__(self, b):
        self.b = b

    def names(self, d=None):
        return ["coRegB"]

    def effective(self, blocked=()):
        return [] if "coRegB" in blocked else ["coRegB"]

    def value(self, X, Y):
        r2 = (X[:, :1] - Y[:, :1].T) ** 2
        return np.exp(-r2 / self.b)

    def grads(self, X, Y):
        r2 = (X[:, :1] - Y[:, :1].T) ** 2
        return {"coRegB": self.value(X, Y) * r2 / abs(self.b) ** 3}


def test_mogp_over_the_product_index_set():
    n, d, q = 90, 3, 3
    X, Y = _problem(n, d, q, 101)
    cov = dg.KroneckerProductKernel(dg.SEKernel(1.5, 1.0), dg.CoRegRBFKernel(2.5), d)
    noise = dg.KroneckerProductKernel(dg.DiracKernel(0.3), dg.CoRegDiracKernel(), d)
    model = dg.MOGPRegressionModel(cov, noise, list(zip(X, Y)), q)
    Z, yflat = orc.mogp_flatten(X, Y)
    ocov = orc.Prod(orc.Sliced(orc.Kernel("se", {"bandwidth": 1.5, "amplitude": 1.0}), 0, d),
                    orc.Sliced(_CoRegRBFOracle(2.5), d))
    # noise: Dirac over x times delta over outputs = plain Dirac on the augmented point
    h = model._current_state
    assert model.energy(h) == pytest.approx(orc.energy(ocov, 0.3, Z, yflat), rel=TOL)
    g, go = model.gradEnergy(h), orc.grad_energy(ocov, 0.3, Z, yflat)
    got = {k.split("/")[-1]: v for k, v in g.items()}
    for name in ("bandwidth", "amplitude", "coRegB"):
        want = [v for k, v in go.items() if k.endswith(name)][0]
        assert got[name] == pytest.approx(want, rel=TOL), name
    assert got["noiseLevel"] == pytest.approx(go["noiseLevel"], rel=TOL)
    Xs, _ = _problem(12, d, q, 102)
    Zs = np.array([dg.MOGPRegressionModel.index(x, o) for x in Xs for o in range(q)])
    post = model.predictiveDistribution(Zs)
    mu, cov_post = orc.predictive(ocov, 0.3, Z, yflat, Zs)
    assert np.abs(post.mu - mu).max() <= TOL * max(1.0, np.abs(mu).max())
    assert np.abs(post.covariance - cov_post).max() <= TOL * np.abs(cov_post).max()


@pytest.mark.parametrize("coreg", ["rbf", "cauchy", "dirac"])
def test_kronecker_mogp_matrix_normal_energy(coreg):
    n, d, q = 150, 3, 4
    X, Y = _problem(n, d, q, 103)
    ck, ock = {"rbf": (dg.CoRegRBFKernel(3.0), _CoRegRBFOracle(3.0)),
               "cauchy": (dg.CoRegCauchyKernel(2.0), orc.Kernel("cauchy", {"sigma": 2.0})),
               "dirac": (dg.CoRegDiracKernel(), orc.dirac(1.0))}[coreg]
    mean = lambda xo: 0.1 * xo[0][0] + 0.05 * xo[1]  # noqa: E731
    model = dg.KroneckerMOGPModel(dg.GenericMaternKernel(1.7, 2), dg.DiracKernel(0.4), ck, list(zip(X, Y)), q, mean)
    M = np.array([[mean((x, o)) for o in range(q)] for x in X])
    want = orc.kronecker_mogp_energy(orc.Kernel("matern", {"p": 2.0, "l": 1.7}), 0.4, ock, X, Y, M)
    assert model.energy(model._current_state) == pytest.approx(want, rel=TOL)
    h = dict(model._current_state)
    lkey = [k for k in h if k.endswith("/l")][0]
    h[lkey] = 2.3
    want2 = orc.kronecker_mogp_energy(orc.Kernel("matern", {"p": 2.0, "l": 2.3}), 0.4, ock, X, Y, M)
    assert model.energy(h) == pytest.approx(want2, rel=TOL)
