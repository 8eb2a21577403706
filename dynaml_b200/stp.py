"""Host-side mirror of the Student-t process regression model: the sibling of
GPRegression that shares the kernel-matrix / Cholesky / solve core and differs only in the scalar
likelihood terms.

Reference: CORE/models/stp/AbstractSTPRegressionModel.scala:34-330,
CORE/probability/distributions/BlockedMultivariateStudentsT.scala:21-90,
CORE/probability/StudentsTRV.scala:65-87 (MultStudentsTPRV).

Reference quirks kept as they are: `setState` stores degrees_of_freedom + 2 (:84), so every
`energy(h)` call evaluates the density at h("degrees_of_freedom") + 2; the log-normaliser uses the
integer division rows/2 (BlockedMultivariateStudentsT.scala:57)."""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from . import _capi
from .kernels import DiracKernel, LocalScalarKernel, _as_points


class BlockedMultivariateStudentsT:
    def __init__(self, mu: float, mean: np.ndarray, covariance: np.ndarray, bars):
        assert mu > 2.0, "Degrees of freedom must be greater than 2.0, for a multivariate t distribution to be defined"
        self.mu, self.mean, self.covariance, self._bars = mu, mean, covariance, bars

    @property
    def variance(self) -> np.ndarray:  # covariance * mu / (mu - 2)   (:66-68)
        return self.covariance * (self.mu / (self.mu - 2.0))

    @property
    def mode(self) -> np.ndarray:
        return self.mean

    def confidenceInterval(self, s: float):
        return self._bars(float(s))


class MultStudentsTPRV:
    def __init__(self, mu: float, mean: np.ndarray, covariance: np.ndarray, bars):
        self.mu, self.mean, self.covariance = mu, mean, covariance
        self.underlyingDist = BlockedMultivariateStudentsT(mu, mean, covariance, bars)


class STPRegression:
    """AbstractSTPRegressionModel(mu, cov, noise, data, num, meanFunc) on one B200."""

    def __init__(self, mu: float, cov: LocalScalarKernel, noise: LocalScalarKernel,
                 trainingdata: Sequence[Tuple[Sequence[float], float]],
                 meanFunc: Callable[[np.ndarray], float] | None = None, ctx: _capi.Context | None = None):
        if not isinstance(noise, DiracKernel):
            raise NotImplementedError("the device path models noise with DiracKernel only; no CPU fallback")
        cov._device_kind()
        self.covariance, self.noiseModel = cov, noise
        self.mean = meanFunc if meanFunc is not None else (lambda _x: 0.0)
        self.g = list(trainingdata)
        self.npoints = len(self.g)
        if self.npoints == 0:
            raise ValueError("training data is empty")
        self.ctx = ctx or _capi.default_context()
        self.hyper_parameters: List[str] = list(cov.hyper_parameters) + list(noise.hyper_parameters) + \
            ["degrees_of_freedom"]
        self.current_state: Dict[str, float] = {**cov.state, **noise.state, "degrees_of_freedom": 2.0 + mu}
        self._data = None

    @property
    def _current_state(self) -> Dict[str, float]:
        return self.current_state

    def setState(self, s: Dict[str, float]):
        # AbstractSTPRegressionModel.scala:73-86 (note the + 2.0)
        self.covariance.setHyperParameters({k: v for k, v in s.items() if k in self.covariance.hyper_parameters})
        self.noiseModel.setHyperParameters({k: v for k, v in s.items() if k in self.noiseModel.hyper_parameters})
        self.current_state = {**self.covariance.state, **self.noiseModel.state}
        self.current_state["degrees_of_freedom"] = s["degrees_of_freedom"] + 2.0
        return self

    def _device_data(self):
        if self._data is None:
            X = _as_points([p[0] for p in self.g])
            y = np.asarray([p[1] for p in self.g], dtype=np.float64)
            m = np.asarray([self.mean(x) for x in X], dtype=np.float64)
            self._data = self.ctx.data_create(X, y, m)
        return self._data

    def __del__(self):
        try:
            if self._data is not None:
                self.ctx.data_destroy(self._data)
        except Exception:
            pass

    def _spec(self):
        return self.covariance._spec(noise=self.noiseModel.state["noiseLevel"])

    @staticmethod
    def neg_log_pdf(mu: float, n: int, quad: float, logdet_half: float) -> float:
        """-(unnormalizedLogPdf - logNormalizer), BlockedMultivariateStudentsT.scala:47-61."""
        unnorm = -0.5 * (mu + n) * math.log(1.0 + quad / mu)
        log_norm = (n // 2) * (math.log(mu) + math.log(math.pi)) + 0.5 * logdet_half + math.lgamma(mu / 2.0) - \
            math.lgamma((mu + n) / 2.0)
        return -1.0 * (unnorm - log_norm)

    def energy(self, h: Dict[str, float], options: Dict[str, str] | None = None) -> float:
        """AbstractSTPRegressionModel.energy (:253-283); +Inf when K is not PD (:296-299)."""
        self.setState(h)
        try:
            quad, logdet = self.ctx.energy_terms(self._device_data(), self._spec())
        except _capi.NotPositiveDefinite:
            return float("inf")
        return self.neg_log_pdf(self.current_state["degrees_of_freedom"], self.npoints, quad, logdet)

    def predictiveDistribution(self, test: Sequence) -> MultStudentsTPRV:
        """:97-188: GP posterior mean; covariance scaled by (nu + beta - 2)/(nu + n - 2), beta = r' K^-1 r."""
        Xs = _as_points(test)
        ms = np.asarray([self.mean(x) for x in Xs], dtype=np.float64)
        model = self.ctx.fit(self._device_data(), self._spec())
        try:
            beta, _ = self.ctx.model_terms(model)
            dof = self.current_state["degrees_of_freedom"]
            adj = (dof + beta - 2.0) / (dof + self.npoints - 2.0)
            mean, cov, lo, hi = self.ctx.predict(model, Xs, ms, want_cov=True, want_bars=True, sigma=1.0)
        finally:
            self.ctx.model_destroy(model)
        mu = self.npoints + dof
        bar1 = (hi - mean) * math.sqrt(adj)  # chol(adj * S) . 1 = sqrt(adj) * chol(S) . 1

        def bars(s: float):
            b = bar1 * (abs(s) * math.sqrt(mu / (mu - 2.0)))  # :79-88
            return mean - b, mean + b

        return MultStudentsTPRV(mu, mean, cov * adj, bars)

    def predictionWithErrorBars(self, testData: Sequence, sigma: int):
        post = self.predictiveDistribution(testData)
        lo, hi = post.underlyingDist.confidenceInterval(float(sigma))
        return [(testData[i], float(post.mean[i]), float(lo[i]), float(hi[i])) for i in range(len(testData))]

    def predict(self, point) -> float:
        return self.predictionWithErrorBars([point], 1)[0][1]


def AbstractSTPRegressionModel(mu, cov, noise, meanFunc=None):
    """AbstractSTPRegressionModel(mu, cov, noise, mean)(data, num) factory (:307-325)."""
    def build(data, num):
        return STPRegression(mu, cov, noise, data[:num] if num > 0 else data, meanFunc)
    return build
