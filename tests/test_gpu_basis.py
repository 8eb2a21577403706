"""GPBasisFuncRegressionModel (CORE/models/gp/GPBasisFuncRegressionModel.scala:70-109):
explicit basis functions for the trend, their prior folded into the kernel matrices as a rank-m update."""
import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def basis(x):
    x = np.asarray(x, dtype=float)
    return np.array([1.0, x[0], x[1], x[0] * x[1], x[2] ** 2])


B = np.array([0.3, -0.2, 0.5, 0.1, -0.4])
COVB = np.diag([0.5, 0.2, 0.3, 0.1, 0.25]) + 0.02


@pytest.mark.parametrize("n", [100, 300])
def test_basis_function_model_matches_oracle(n):
    X, y, Xs = orc.synthetic(n, 3, seed=61, m=45)
    k, ok = dg.SEKernel(1.6, 1.1), orc.Kernel("se", {"bandwidth": 1.6, "amplitude": 1.1})
    model = dg.GPBasisFuncRegression(k, dg.DiracKernel(0.3), list(zip(X, y)), basis, (B, COVB))
    ofm = orc.WithFeatureMap(ok, basis, COVB)
    mean = np.array([basis(x) @ B for x in X])
    mean_s = np.array([basis(x) @ B for x in Xs])
    h = model._current_state
    assert model.energy(h) == pytest.approx(orc.energy(ofm, 0.3, X, y, mean), rel=TOL)
    # gradEnergy is inherited: covariance + noise without the feature-map term (AbstractGPRegressionModel.scala:196-253)
    g, go = model.gradEnergy(h), orc.grad_energy(ok, 0.3, X, y, mean)
    for name, v in go.items():
        assert g[name] == pytest.approx(v, rel=TOL), name
    post = model.predictiveDistribution(Xs)
    mu, cov = orc.predictive(ofm, 0.3, X, y, Xs, mean, mean_s)
    assert np.abs(post.mu - mu).max() <= TOL * max(1.0, np.abs(mu).max())
    assert np.abs(post.covariance - cov).max() <= TOL * np.abs(cov).max()
    bars = model.predictionWithErrorBars(list(Xs), 2)
    lo, hi = orc.error_bars(mu, cov, 2.0)
    assert np.abs(np.array([b_[2] for b_ in bars]) - lo).max() <= 1e-7 * np.abs(lo).max()
    # energy after a state change re-uses the resident features
    h2 = dict(h)
    h2["bandwidth"] = 2.2
    ok2 = orc.WithFeatureMap(orc.Kernel("se", {"bandwidth": 2.2, "amplitude": 1.1}), basis, COVB)
    assert model.energy(h2) == pytest.approx(orc.energy(ok2, 0.3, X, y, mean), rel=TOL)


def test_basis_gradient_with_feature_map_term():
    """grad_with_basis=True differentiates the matrix the energy uses (the trend prior included)."""
    X, y, _ = orc.synthetic(150, 3, seed=63)
    k, ok = dg.SEKernel(1.6, 1.1), orc.Kernel("se", {"bandwidth": 1.6, "amplitude": 1.1})
    model = dg.GPBasisFuncRegression(k, dg.DiracKernel(0.3), list(zip(X, y)), basis, (B, COVB), grad_with_basis=True)
    mean = np.array([basis(x) @ B for x in X])
    g = model.gradEnergy(model._current_state)
    go = orc.grad_energy(orc.WithFeatureMap(ok, basis, COVB), 0.3, X, y, mean)
    for name, v in go.items():
        assert g[name] == pytest.approx(v, rel=TOL), name


def test_predict_without_test_features_is_an_error(ctx):
    X, y, Xs = orc.synthetic(80, 3, seed=65, m=10)
    data = ctx.data_create(X, y)
    ctx.data_set_basis(data, np.array([basis(x) for x in X]))
    model = ctx.fit(data, dg.SEKernel(1.5, 1.0)._spec(noise=0.2))
    try:
        with pytest.raises(dg.DgpError):
            ctx.predict(model, Xs, None, want_cov=False, want_bars=False)
        with pytest.raises(dg.DgpError):
            ctx.model_set_test_basis(model, np.ones((10, 3)))  # wrong number of basis functions
        ctx.model_set_test_basis(model, np.array([basis(x) for x in Xs]))
        mean, _, _, _ = ctx.predict(model, Xs, None, want_cov=False, want_bars=False)
        assert np.isfinite(mean).all()
        ctx.data_set_basis(data, None)   # removing the term gives the plain GP energy again
        e = ctx.energy(data, dg.SEKernel(1.5, 1.0)._spec(noise=0.2))
        assert e == pytest.approx(orc.energy(orc.Kernel("se", {"bandwidth": 1.5, "amplitude": 1.0}), 0.2, X, y), rel=TOL)
    finally:
        ctx.model_destroy(model)
        ctx.data_destroy(data)
