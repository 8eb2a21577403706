"""Student-t process regression on the device terms of the GP path, against the
reference's STPSpec known answers and the oracle."""
import math

import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_stpspec_likelihood():
    # STPSpec.scala:13-39
    data = [([0.0], 0.0), ([1.0], 1.0)]
    model = dg.AbstractSTPRegressionModel(4.0, dg.DiracKernel(0.5), dg.DiracKernel(0.5), lambda v: v[0])(data, 2)
    likelihood = model.energy(model._current_state)
    assert abs(likelihood - (-math.log(1 / (2.0 * math.pi)))) < 1e-5


def test_stpspec_posterior_and_error_bars():
    # STPSpec.scala:41-80
    data = [([1.0], 1.0), ([2.0], 4.0)]
    noise = dg.DiracKernel(0.0)
    noise.block_all_hyper_parameters()
    model = dg.AbstractSTPRegressionModel(4.0, dg.SEKernel(1.0, 1.0), noise, lambda v: 0.0)(data[:1], 1)
    post = model.predictiveDistribution([data[1][0]])
    rho = math.exp(-0.5)
    assert post.underlyingDist.mean[0] == pytest.approx(rho, rel=4e-16)
    assert abs(post.underlyingDist.variance[0, 0] - (1.0 - rho * rho) * 7 / 5) < 1e-5
    bars = model.predictionWithErrorBars([data[1][0]], 1)
    assert bars[0][2] == pytest.approx(rho - math.sqrt((1.0 - rho * rho) * 7 / 5), abs=1e-14)
    assert bars[0][3] == pytest.approx(rho + math.sqrt((1.0 - rho * rho) * 7 / 5), abs=1e-14)


@pytest.mark.parametrize("n", [50, 301, 1200])
def test_stp_energy_and_prediction_vs_oracle(n):
    X, y, Xs = orc.synthetic(n, 4, seed=n, m=30)
    k, ok = dg.GenericMaternKernel(2.0, 2), orc.Kernel("matern", {"p": 2.0, "l": 2.0})
    model = dg.STPRegression(5.0, k, dg.DiracKernel(0.4), list(zip(X, y)))
    h = dict(model._current_state)
    e = model.energy(h)  # evaluates at h["degrees_of_freedom"] + 2 (reference quirk)
    assert e == pytest.approx(orc.stp_energy(h["degrees_of_freedom"] + 2.0, ok, 0.4, X, y), rel=TOL)
    dof = model._current_state["degrees_of_freedom"]
    post = model.predictiveDistribution(Xs)
    mu, mean, cov = orc.stp_predictive(dof, ok, 0.4, X, y, Xs)
    assert post.mu == mu
    assert np.abs(post.mean - mean).max() <= TOL * np.abs(mean).max()
    assert np.abs(post.covariance - cov).max() <= TOL * np.abs(cov).max()
    lo, hi = post.underlyingDist.confidenceInterval(2.0)
    lo2, hi2 = orc.stp_error_bars(mu, mean, cov, 2.0)
    assert np.abs(lo - lo2).max() <= 1e-8 * np.abs(lo2).max() and np.abs(hi - hi2).max() <= 1e-8 * np.abs(hi2).max()


def test_stp_non_pd_is_inf():
    X, y, _ = orc.synthetic(100, 2, seed=2)
    model = dg.STPRegression(4.0, dg.RBFKernel(1.0), dg.DiracKernel(0.3), list(zip(X, y)))
    assert model.energy({"bandwidth": float("nan"), "noiseLevel": 0.3, "degrees_of_freedom": 4.0}) == float("inf")
