"""Hyper-parameter samplers that consume `energy`:
AdaptiveHyperParameterMCMC, HyperParameterMCMC and HyperParameterSCAM of
CORE/probability/mcmc/HyperParameterMCMC.scala:49-345.  The chain logic is host code, as in the reference; every
log-likelihood evaluation is one resident-data `energy` call on the device (X, y and the workspaces stay in HBM,
so a step costs one kernel-matrix assembly + Cholesky + solves and a scalar read-back).

Priors are objects with `.logPdf(x)` (the mean-field prior of the reference is their sum); randomness comes from
an injected numpy Generator so that chains are reproducible."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

Config = Dict[str, float]


class ConfigEncoding:
    """utils.ConfigEncoding: a fixed key order between Map[String, Double] and DenseVector[Double]."""

    def __init__(self, keys: Sequence[str]):
        self.keys = list(keys)

    def __call__(self, config: Config) -> np.ndarray:
        return np.array([float(config[k]) for k in self.keys])

    def i(self, v: np.ndarray) -> Config:
        return {k: float(x) for k, x in zip(self.keys, v)}


class AdaptiveHyperParameterMCMC:
    """Adaptive Metropolis (:133-310): Gaussian proposal around the previous sample whose covariance follows the
    running covariance of the chain."""

    algoName = "Adaptive MCMC"

    def __init__(self, system, hyper_prior: Dict[str, object], burnIn: int, rng: np.random.Generator | None = None):
        self.system, self.hyper_prior, self.burnIn = system, dict(hyper_prior), int(burnIn)
        self.rng = rng if rng is not None else np.random.default_rng()
        self.encoder = ConfigEncoding(list(self.hyper_prior.keys()))
        self.initialState = self.encoder(system._current_state)
        self.dimensions = len(self.hyper_prior)
        self.beta = 0.05
        self.count = 1
        self.sigma = np.outer(self.initialState, self.initialState)
        self.acceptedSamples = 0
        self.mean = self.initialState.copy()
        self.previous_sample = self.initialState.copy()
        self.previous_log_likelihood = self.logLikelihood(self.initialState)
        self.eye = np.eye(self.dimensions)
        for _ in range(self.burnIn):
            self.next()

    # log prior - energy (:158-163)
    def logLikelihood(self, candidate: np.ndarray) -> float:
        cfg = self.encoder.i(candidate)
        prior = sum(self.hyper_prior[k].logPdf(v) for k, v in cfg.items())
        return prior - self.system.energy(cfg)

    def candidateDistribution(self, mean: np.ndarray, covariance: np.ndarray):
        """MultGaussianRV(mean, covariance): draw = mean + cholesky(covariance) z."""
        rng = self.rng

        class _Gaussian:
            def draw(self_inner) -> np.ndarray:
                L = np.linalg.cholesky(covariance)
                return mean + L @ rng.standard_normal(mean.size)
        return _Gaussian()

    @property
    def _previous_sample(self) -> Config:
        return self.encoder.i(self.previous_sample)

    @property
    def _previous_log_likelihood(self) -> float:
        return self.previous_log_likelihood

    @property
    def sampleAcceptenceRate(self) -> float:
        return self.acceptedSamples / self.count

    @property
    def _count(self) -> int:
        return self.count

    def getExplorationVar(self) -> np.ndarray:
        """:221-227."""
        d = float(self.dimensions)
        if self.count <= 2 * self.dimensions:
            return self.eye * 0.01 / d
        adj = self.count / (self.count - 1)
        return self.sigma * ((1 - self.beta) * 2.38) ** 2 * adj / d + self.eye * (0.1 * self.beta) ** 2 / d

    def next(self) -> np.ndarray:
        """:229-276: propose until one candidate is accepted; the moments see every proposal round."""
        explorationCov = self.getExplorationVar()
        accepted = None
        while accepted is None:
            self.count += 1
            candidate = self.candidateDistribution(self.previous_sample, explorationCov).draw()
            likelihood = self.logLikelihood(candidate)
            ratio = likelihood - self.previous_log_likelihood
            if ratio > 0.0 or float(self.rng.random()) < math.exp(ratio):
                self.previous_sample = candidate
                self.previous_log_likelihood = likelihood
                self.acceptedSamples += 1
                accepted = candidate
                x = candidate
            else:
                x = self.previous_sample
            self.updateMoments(x)
        return accepted

    def updateMoments(self, x: np.ndarray) -> None:
        """:278-287."""
        m, s = self.mean, self.sigma
        mnew = m + (x - m) / (self.count + 1)
        snew = s + np.outer(m, m) - np.outer(mnew, mnew) + (np.outer(x, x) - s - np.outer(m, m)) / (self.count + 1)
        self.mean, self.sigma = mnew, snew

    def sample(self) -> Config:
        return self.encoder.i(self.next())

    draw = sample

    def iid(self, n: int) -> List[Config]:
        return [self.encoder.i(self.next()) for _ in range(n)]


class HyperParameterMCMC(AdaptiveHyperParameterMCMC):
    """Vanilla Metropolis-Hastings (:49-76): candidates are draws of the fixed `proposal` random variable
    (an object with `.draw()`), the exploration covariance is never adapted."""

    algoName = "Vanilla MCMC"

    def __init__(self, system, hyper_prior, proposal, burnIn: int, rng=None):
        self.proposal = proposal
        super().__init__(system, hyper_prior, burnIn, rng)

    def candidateDistribution(self, mean, covariance):
        return self.proposal

    def getExplorationVar(self):
        return self.sigma

    def updateMoments(self, x):
        pass


class HyperParameterSCAM(AdaptiveHyperParameterMCMC):
    """Single Component Adaptive Metropolis (:324-345): diagonal exploration covariance."""

    algoName = "Hyper-Parameter SCAM"

    def getExplorationVar(self):
        if self.count <= 10:
            return self.eye * 25.0
        adj = self.count / (self.count - 1)
        return np.diag((np.diag(self.sigma) * adj + 0.05) * 2.4 * 2.4)
