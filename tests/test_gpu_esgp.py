"""ESGPModel (extended skew GP) and the two building blocks it rides on, dgp_model_solve and
dgp_model_cross_apply, against the oracle."""
import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_model_solve_and_cross_apply(ctx):
    X, y, Xs = orc.synthetic(300, 3, seed=81, m=70)
    k, ok = dg.GenericMaternKernel(1.8, 2), orc.Kernel("matern", {"p": 2.0, "l": 1.8})
    data = ctx.data_create(X, y)
    model = ctx.fit(data, k._spec(noise=0.3))
    try:
        K = orc.build_kernel_matrix(ok, X, orc.dirac(0.3))
        B = np.random.default_rng(1).standard_normal((300, 3))
        got = ctx.model_solve(model, B)
        want = np.linalg.solve(K, B)
        assert np.abs(got - want).max() <= TOL * np.abs(want).max()
        assert np.abs(ctx.model_solve(model, B[:, 0]) - want[:, 0]).max() <= TOL * np.abs(want).max()
        Ks = orc.cross_kernel_matrix(ok, X, Xs)
        got = ctx.model_cross_apply(model, Xs, B)
        assert np.abs(got - Ks.T @ B).max() <= 1e-12 * np.abs(Ks.T @ B).max()
    finally:
        ctx.model_destroy(model)
        ctx.data_destroy(data)


@pytest.mark.parametrize("lam,tau", [(0.0, 0.2), (0.4, 0.2), (-0.3, -0.5), (1.2, 1.0)])
def test_esgp_energy_matches_oracle(lam, tau):
    X, y, _ = orc.synthetic(260, 3, seed=83)
    mean = lambda x: 0.2 * x[0]  # noqa: E731
    m = 0.2 * X[:, 0]
    k, ok = dg.SEKernel(1.5, 1.1), orc.Kernel("se", {"bandwidth": 1.5, "amplitude": 1.1})
    model = dg.ESGPModel(k, dg.DiracKernel(0.3), list(zip(X, y)), lam, tau, mean)
    assert model._hyper_parameters == ["bandwidth", "amplitude", "noiseLevel", "skewness", "cutoff"]
    e = model.energy(model._current_state)
    assert e == pytest.approx(orc.esgp_energy(ok, 0.3, X, y, lam, tau, m), rel=TOL)
    if lam == 0.0:  # without skewness the model is the GP
        gp = dg.GPRegression(dg.SEKernel(1.5, 1.1), dg.DiracKernel(0.3), list(zip(X, y)), mean)
        assert e == pytest.approx(gp.energy(gp._current_state), rel=1e-12)
    h = dict(model._current_state)
    h.update(bandwidth=2.1, skewness=lam + 0.1, cutoff=tau - 0.1)
    ok2 = orc.Kernel("se", {"bandwidth": 2.1, "amplitude": 1.1})
    assert model.energy(h) == pytest.approx(orc.esgp_energy(ok2, 0.3, X, y, lam + 0.1, tau - 0.1, m), rel=TOL)


def test_esgp_composite_covariance_and_predictive():
    X, y, Xs = orc.synthetic(220, 2, seed=85, m=30)
    k = dg.SEKernel(1.4, 1.0) + dg.CauchyKernel(2.5) * 0.5
    ok = orc.Sum(orc.Kernel("se", {"bandwidth": 1.4, "amplitude": 1.0}), orc.Scaled(orc.Kernel("cauchy", {"sigma": 2.5}), 0.5))
    model = dg.ESGPModel(k, dg.DiracKernel(0.25), list(zip(X, y)), 0.6, 0.3)
    assert model.energy(model._current_state) == pytest.approx(orc.esgp_energy(ok, 0.25, X, y, 0.6, 0.3), rel=TOL)
    post = model.predictiveDistribution(Xs)
    mu, cov, skew, cutoff = orc.esgp_predictive(ok, 0.25, X, y, Xs, 0.6, 0.3)
    assert np.abs(post.mu - mu).max() <= TOL * max(1.0, np.abs(mu).max())
    assert np.abs(post.sigma - cov).max() <= TOL * np.abs(cov).max()
    assert np.abs(post.alpha - skew).max() <= TOL * np.abs(skew).max()
    assert post.tau == pytest.approx(cutoff, rel=TOL)


def test_esgp_non_pd_is_infinite():
    X, y, _ = orc.synthetic(100, 2, seed=87)
    model = dg.ESGPModel(dg.SEKernel(1.5, 1.0), dg.DiracKernel(0.3), list(zip(X, y)), 0.5, 0.1)
    h = dict(model._current_state)
    h["bandwidth"] = float("nan")
    assert model.energy(h) == float("inf")
