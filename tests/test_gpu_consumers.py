"""Consumers of `energy`: the GP committee of ProbGPCommMachine and the hyper-parameter
samplers, driving the device path."""
import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_committee_gp_is_the_weighted_sum_of_its_members():
    X, y, Xs = orc.synthetic(240, 3, seed=71, m=25)
    model = dg.GPRegression(dg.SEKernel(1.5, 1.0), dg.DiracKernel(0.3), list(zip(X, y)))
    model.covariance.block("amplitude")
    machine = dg.ProbGPCommMachine(model, rng=np.random.default_rng(5))
    machine.setPolicy("GS").setBaseLinePolicy("mean").setGridSize(3).setStepSize(0.25)
    committee, state = machine.optimize({"bandwidth": 2.2, "noiseLevel": 0.6}, {})
    weights = machine.weights
    assert len(weights) == 9 and all(w > 0 for w, _ in weights)
    # oracle: sum_i w_i^2 SE(bandwidth_i) with noise sum_i w_i^2 |noise_i| and the (zero) mean scaled by sum w
    ok = None
    noise = 0.0
    for i, (w, cfg) in enumerate(weights):
        member = orc.Scaled(orc.Kernel("se", {"bandwidth": cfg["bandwidth"], "amplitude": 1.0}), w * w)
        ok = member if ok is None else orc.Sum(ok, member, ids=(f"s{i}", f"m{i}"))
        noise += w * w * abs(cfg["noiseLevel"])
    assert committee.noiseModel.state["noiseLevel"] == pytest.approx(noise, rel=1e-15)
    assert committee.energy(state) == pytest.approx(orc.energy(ok, noise, X, y), rel=TOL)
    post = committee.predictiveDistribution(Xs)
    mu, cov = orc.predictive(ok, noise, X, y, Xs)
    assert np.abs(post.mu - mu).max() <= TOL * max(1.0, np.abs(mu).max())
    assert np.abs(post.covariance - cov).max() <= TOL * np.abs(cov).max()
    # the committee's own gradient: one entry per member hyper-parameter (27 base-kernel parameters) + noise
    g = committee.gradEnergy(state)
    go = orc.grad_energy(ok, noise, X, y)
    ref = [v for n_, v in go.items() if n_.endswith("bandwidth")]
    got = [v for n_, v in g.items() if n_.endswith("bandwidth")]
    assert len(got) == 9 and np.allclose(got, ref, rtol=TOL, atol=TOL * max(abs(v) for v in ref))


def test_csa_committee_runs_on_the_device():
    X, y, _ = orc.synthetic(200, 2, seed=73)
    model = dg.GPRegression(dg.RBFKernel(1.2), dg.DiracKernel(0.4), list(zip(X, y)))
    machine = dg.ProbGPCommMachine(model, rng=np.random.default_rng(9))
    machine.setGridSize(2).setStepSize(0.2).setMaxIterations(3)
    committee, state = machine.optimize({"bandwidth": 1.8, "noiseLevel": 0.7}, {"persist": "true"})
    assert len(machine.weights) == 4 and np.isfinite(committee.energy(state))
    assert np.isfinite(committee.predict(X[0]))


class _LogNormalPrior:
    def __init__(self, mu, s):
        self.mu, self.s = mu, s

    def logPdf(self, x):
        if x <= 0:
            return -np.inf
        z = (np.log(x) - self.mu) / self.s
        return float(-0.5 * z * z - np.log(x * self.s * np.sqrt(2 * np.pi)))


def test_adaptive_mcmc_over_a_device_gp():
    X, y, _ = orc.synthetic(150, 2, seed=75)
    model = dg.GPRegression(dg.RBFKernel(1.5), dg.DiracKernel(0.3), list(zip(X, y)))
    prior = {"bandwidth": _LogNormalPrior(0.3, 0.6), "noiseLevel": _LogNormalPrior(-1.0, 0.6)}
    mc = dg.AdaptiveHyperParameterMCMC(model, prior, burnIn=10, rng=np.random.default_rng(11))
    samples = mc.iid(15)
    assert len(samples) == 15 and all(s["bandwidth"] > 0 and s["noiseLevel"] > 0 for s in samples)
    last = mc._previous_sample
    ok = orc.Kernel("rbf", {"bandwidth": last["bandwidth"]})
    want = sum(prior[k].logPdf(v) for k, v in last.items()) - orc.energy(ok, last["noiseLevel"], X, y)
    assert mc._previous_log_likelihood == pytest.approx(want, rel=TOL)
