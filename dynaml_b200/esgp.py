"""ESGPModel -- extended skew Gaussian process regression (CORE/models/sgp/ESGPModel.scala:40-430,
CORE/probability/distributions/SkewGaussian.scala:206-262, SkewSymmDistribution.scala:27-62).

The sibling shares the blocked Cholesky / solve core of the GP path: with Sigma = K_cov + K_noise, skewness vector
alpha = lambda 1, cutoff tau and trend mu, the negative log-likelihood needs the Cholesky factors of Sigma and of
Sigma + alpha alpha' (a constant kernel lambda^2 added to the covariance expression, i.e. a device composite) and
Sigma^-1 applied to two vectors; the posterior predictive needs the GP posterior plus Sigma^-1 alpha and k(X, X*)' of
it.  Everything O(N^2) and above runs on the device (dgp_fit, dgp_energy_terms, dgp_model_solve, dgp_predict,
dgp_model_cross_apply); the scalar glue below is the reference's own."""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from . import _capi
from .kernels import DiracKernel, LocalScalarKernel, _as_points

LOG_2PI = math.log(2.0 * math.pi)


def _std_normal_cdf(x: float) -> float:
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))  # breeze Gaussian(0, 1).cdf


class BlockedMESNRV:
    """Parameters of a multivariate extended skew normal (probability/BlockedMESNRV.scala:27)."""

    def __init__(self, tau: float, alpha: np.ndarray, mu: np.ndarray, sigma: np.ndarray):
        self.tau, self.alpha, self.mu, self.sigma = float(tau), alpha, mu, sigma


class ESGPModel:
    def __init__(self, cov: LocalScalarKernel, noise: LocalScalarKernel, trainingdata: Sequence[Tuple[Sequence[float], float]],
                 lambda_: float, tau: float, meanFunc: Callable[[np.ndarray], float] | None = None,
                 ctx: _capi.Context | None = None):
        if not isinstance(noise, DiracKernel):
            raise NotImplementedError("the device path models noise with DiracKernel only; no CPU fallback")
        cov._device_kind()
        self.covariance, self.noiseModel = cov, noise
        self.mean = meanFunc if meanFunc is not None else (lambda _x: 0.0)
        self.g = list(trainingdata)
        if not self.g:
            raise ValueError("training data is empty")
        self.npoints = len(self.g)
        self.ctx = ctx or _capi.default_context()
        self.hyper_parameters: List[str] = list(cov.hyper_parameters) + list(noise.hyper_parameters) + ["skewness", "cutoff"]
        self.current_state: Dict[str, float] = {**cov.state, **noise.state, "skewness": float(lambda_), "cutoff": float(tau)}
        self._X = _as_points([p[0] for p in self.g])
        self._y = np.asarray([p[1] for p in self.g], dtype=np.float64)
        self._mu = np.asarray([self.mean(x) for x in self._X], dtype=np.float64)
        self._handles: Dict[float, object] = {}

    @property
    def _current_state(self) -> Dict[str, float]:
        return self.current_state

    @property
    def _hyper_parameters(self) -> List[str]:
        return self.hyper_parameters

    def setState(self, s: Dict[str, float]):
        # ESGPModel.scala:94-108
        self.covariance.setHyperParameters({k: v for k, v in s.items() if k in self.covariance.hyper_parameters})
        self.noiseModel.setHyperParameters({k: v for k, v in s.items() if k in self.noiseModel.hyper_parameters})
        self.current_state = {**self.covariance.state, **self.noiseModel.state,
                              "skewness": float(s["skewness"]), "cutoff": float(s["cutoff"])}
        return self

    def __del__(self):
        try:
            for h in self._handles.values():
                self.ctx.data_destroy(h)
        except Exception:
            pass

    # ---- device residency: one data handle per centre shift (y - mu - shift is what the solves see) ----------
    def _data(self, shift: float):
        if shift not in self._handles:
            if len(self._handles) >= 4:
                k0 = next(iter(self._handles))
                self.ctx.data_destroy(self._handles.pop(k0))
            self._handles[shift] = self.ctx.data_create(self._X, self._y, self._mu + shift)
        return self._handles[shift]

    def _spec_sigma(self):
        return self.covariance._spec(noise=self.noiseModel.state["noiseLevel"])

    def _spec_adjusted(self, lam: float):
        """covariance + noise + alpha alpha' with alpha = lambda 1: the constant lambda^2 is one more term (all
        exponents zero) of the covariance's sum-of-products form."""
        from .kernels import composite_spec
        return composite_spec([(self.covariance, self.covariance.state)], self.noiseModel.state["noiseLevel"],
                              extra_terms=[(lam * lam, {})])

    # ---- GloballyOptimizable ----------------------------------------------------------------------------------
    def energy(self, h: Dict[str, float], options: Dict[str, str] | None = None) -> float:
        """-logPdf of BlockedMESN(tau, alpha, mu, Sigma) at the training labels (ESGPModel.scala:246-273, 337-348):
        basis N(mu + alpha tau, Sigma + alpha alpha'), warping Phi(alpha' Sigma^-1 (y - centre) / delta + tau delta),
        delta = sqrt(1 + alpha' Sigma^-1 alpha); +Inf when a Cholesky fails."""
        self.setState(h)
        lam, tau = self.current_state["skewness"], self.current_state["cutoff"]
        n = self.npoints
        data = self._data(lam * tau)                    # residual y - (mu + alpha tau)
        try:
            quad_v, logsum_v = self.ctx.energy_terms(data, self._spec_adjusted(lam))
            model = self.ctx.fit(data, self._spec_sigma())
        except _capi.NotPositiveDefinite:
            return float("inf")
        try:
            r = self._y - self._mu - lam * tau
            sol = self.ctx.model_solve(model, np.column_stack([np.ones(n), r]))
        finally:
            self.ctx.model_destroy(model)
        delta = math.sqrt(1.0 + lam * lam * float(sol[:, 0].sum()))
        w = lam * float(sol[:, 1].sum()) / delta
        unnorm = -0.5 * quad_v + math.log(_std_normal_cdf(w + tau * delta))
        log_norm = math.log(_std_normal_cdf(tau)) + 0.5 * n * LOG_2PI + 0.5 * logsum_v
        return -(unnorm - log_norm)

    # ---- prediction (ESGPModel.scala:133-198, solve :349-389) ---------------------------------------------------
    def predictiveDistribution(self, test: Sequence) -> BlockedMESNRV:
        lam, tau = self.current_state["skewness"], self.current_state["cutoff"]
        Xs = _as_points(test)
        ms = np.asarray([self.mean(x) for x in Xs], dtype=np.float64)
        n = self.npoints
        model = self.ctx.fit(self._data(0.0), self._spec_sigma())
        try:
            mean, cov, _, _ = self.ctx.predict(model, Xs, ms, want_cov=True, want_bars=False)
            sol = self.ctx.model_solve(model, np.column_stack([self._y - self._mu, np.full(n, lam)]))
            alpha_vec, beta = sol[:, 0], sol[:, 1]
            cross_beta = self.ctx.model_cross_apply(model, Xs, beta)
        finally:
            self.ctx.model_destroy(model)
        delta = 1.0 / math.sqrt(1.0 + lam * float(beta.sum()))
        skew = (np.full(Xs.shape[0], lam) - cross_beta) * delta
        cutoff = (tau + lam * float(alpha_vec.sum())) * delta
        return BlockedMESNRV(cutoff, skew, mean, cov)
