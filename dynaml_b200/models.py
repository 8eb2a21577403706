"""Host-side mirror of AbstractGPRegressionModel / GPRegression and their result types.

Reference: CORE/models/gp/AbstractGPRegressionModel.scala:60-535, CORE/models/gp/GPRegression.scala:51-175,
CORE/probability/GaussianRV.scala:81-103 (MultGaussianPRV),
CORE/probability/distributions/BlockedMultiVariateGaussian.scala:21-68.

`CudaGPRegression` is the class INTEGRATION.md describes on the JVM side: it overrides energy,
gradEnergy, predictiveDistribution, predictionWithErrorBars, persist / unpersist and forwards
them to the C ABI.  `GPRegression` is an alias so reference-style code reads unchanged.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from . import _capi
from .kernels import DiracKernel, LocalScalarKernel, _as_points


class BlockedMultiVariateGaussian:
    """mean / covariance plus the reference's error-bar rule (root * ones, lines 58-68)."""

    def __init__(self, mean: np.ndarray, covariance: np.ndarray, _bars: Callable[[float], Tuple[np.ndarray, np.ndarray]]):
        self.mean = mean
        self.covariance = covariance
        self._bars = _bars

    @property
    def variance(self):
        return self.covariance

    @property
    def mode(self):
        return self.mean

    def confidenceInterval(self, s: float):
        return self._bars(float(s))


class MultGaussianPRV:
    def __init__(self, mu: np.ndarray, covariance: np.ndarray, bars):
        self.mu = mu
        self.covariance = covariance
        self.underlyingDist = BlockedMultiVariateGaussian(mu, covariance, bars)


class CudaGPRegression:
    """GPRegression(cov, noise, trainingdata, meanFunc) on one B200."""

    def __init__(self, cov: LocalScalarKernel, noise: LocalScalarKernel | None = None,
                 trainingdata: Sequence[Tuple[Sequence[float], float]] = (),
                 meanFunc: Callable[[np.ndarray], float] | None = None,
                 ctx: _capi.Context | None = None):
        if noise is None:
            noise = DiracKernel(1.0)
        cov._device_kind()  # raises for unsupported covariance functions
        noise._device_kind()
        # a DiracKernel noise is the spec's dedicated noise term; any other LocalScalarKernel joins the training
        # matrix as part of a composite and is left out of the cross / test matrices (:269-295)
        self._dirac_noise = isinstance(noise, DiracKernel)
        self.covariance, self.noiseModel = cov, noise
        self.mean = meanFunc if meanFunc is not None else (lambda _x: 0.0)
        self.g = list(trainingdata)
        self.npoints = len(self.g)
        if self.npoints == 0:
            raise ValueError("training data is empty")
        self.ctx = ctx or _capi.default_context()
        self.blockSize = 1000
        self.hyper_parameters: List[str] = list(cov.hyper_parameters) + list(noise.hyper_parameters)
        self.current_state: Dict[str, float] = {**cov.state, **noise.state}
        self.validationSet: List[Tuple[Sequence[float], float]] = []
        self._data = None        # dgp_data handle (training [+ validation] set resident in HBM)
        self._data_key = None
        self._model = None       # dgp_model handle (persisted L, alpha)
        self._train_data = None  # second resident copy without the validation set (persist / predict), made lazily
        self._peers: List[Tuple[_capi.Context, object]] = []  # (context, data handle) on further GPUs for energyBatch
        self.caching = False

    # ---- reference accessors ----------------------------------------------------------------
    def blockSize_(self, b: int) -> None:
        self.blockSize = int(b)  # kept for API parity; the device path tiles on its own

    @property
    def _blockSize(self) -> int:
        return self.blockSize

    @property
    def _current_state(self) -> Dict[str, float]:
        return self.current_state

    @property
    def _hyper_parameters(self) -> List[str]:
        return self.hyper_parameters

    def validationSet_(self, v, append: bool = False) -> None:
        self.validationSet = (self.validationSet + list(v)) if append else list(v)
        self._drop_data()

    def dataAsSeq(self, data):
        return data

    def setState(self, s: Dict[str, float]):
        # AbstractGPRegressionModel.scala:117-128
        self.covariance.setHyperParameters({k: v for k, v in s.items() if k in self.covariance.hyper_parameters})
        self.noiseModel.setHyperParameters({k: v for k, v in s.items() if k in self.noiseModel.hyper_parameters})
        self.current_state = {**self.covariance.state, **self.noiseModel.state}
        return self

    # ---- further GPUs for the candidate loop (one process) ---------------------------
    def useDevices(self, devices: Sequence[int]):
        """Shard `energyBatch` (hence GridSearch / CSA sweeps) over these GPUs from this one process: the training
        set is replicated on each, candidate i runs on devices[i % len(devices)] (dgp_energy_batch_multi).  The
        model's own context counts as the first device; energy / gradEnergy / prediction stay on it."""
        self._drop_peers()
        own_seen = False
        for dev in devices:
            if int(dev) == self.ctx.device and not own_seen:
                own_seen = True      # the model's own context
                continue
            self._peers.append((_capi.Context(int(dev)), None))   # a further context (a device may be listed twice)
        return self

    def _drop_peers(self, contexts: bool = True):
        for i, (c, h) in enumerate(self._peers):
            if h is not None:
                c.data_destroy(h)
            self._peers[i] = (c, None)
        if contexts:
            for c, _ in self._peers:
                c.close()
            self._peers = []

    # ---- device residency -------------------------------------------------------------------
    def _drop_data(self):
        self._drop_peers(contexts=False)
        if self._data is not None:
            self.ctx.data_destroy(self._data)
            self._data = None
        if getattr(self, "_train_data", None) is not None:
            self.ctx.data_destroy(self._train_data)
            self._train_data = None

    def _drop_model(self):
        if self._model is not None:
            self.ctx.model_destroy(self._model)
            self._model = None

    def __del__(self):
        try:
            self._drop_model()
            self._drop_data()
            self._drop_peers()
        except Exception:
            pass

    def _device_data(self):
        """Training (+ validation, GPRegression.scala:148-174) set packed once into HBM."""
        if self._data is None:
            pts = self.g + self.validationSet
            X = _as_points([p[0] for p in pts])
            y = np.asarray([p[1] for p in pts], dtype=np.float64)
            m = np.asarray([self.mean(x) for x in X], dtype=np.float64)
            self._data = self.ctx.data_create(X, y, m)
            self._attach_basis(self._data, X)
        return self._data

    # hooks of GPBasisFuncRegression (explicit basis functions); no-ops for the plain model
    def _attach_basis(self, handle, X: np.ndarray) -> None:
        pass

    def _attach_test_basis(self, model, Xs: np.ndarray) -> None:
        pass

    def _attach_basis_on(self, ctx, handle, X: np.ndarray) -> None:
        pass

    def _spec(self):
        if self._dirac_noise:
            from .kernels import AdditiveCovariance
            if isinstance(self.covariance, AdditiveCovariance):
                return self.covariance._spec(noise=self.noiseModel.state["noiseLevel"], as_model_covariance=True)
            return self.covariance._spec(noise=self.noiseModel.state["noiseLevel"])
        from .kernels import composite_spec
        return composite_spec([(self.covariance, self.covariance.state), (self.noiseModel, self.noiseModel.state)], 0.0)

    def _fit(self, data):
        model = self.ctx.fit(data, self._spec())
        if not self._dirac_noise:
            try:
                self.ctx.model_set_cross_spec(model, self.covariance._spec())
            except Exception:
                self.ctx.model_destroy(model)
                raise
        return model

    # ---- GloballyOptimizable ----------------------------------------------------------------
    def energy(self, h: Dict[str, float], options: Dict[str, str] | None = None) -> float:
        """0.5*(r'K^-1 r + sum log L_ii + n log 2pi); +Inf when K is not PD (:148-189, 516-535)."""
        self.setState(h)
        return self.ctx.energy(self._device_data(), self._spec())

    def energyBatch(self, hs: Sequence[Dict[str, float]], options: Dict[str, str] | None = None) -> List[float]:
        """The serial candidate loop of ModelTuner.getEnergyLandscape as one device batch.
        Leaves the model state at the last candidate, as the serial loop would."""
        specs = []
        for h in hs:
            self.setState(h)
            specs.append(self._spec())
        if not specs:
            return []
        if self._peers:
            pts = self.g + self.validationSet
            for i, (c, h) in enumerate(self._peers):
                if h is None:
                    X = _as_points([p[0] for p in pts])
                    y = np.asarray([p[1] for p in pts], dtype=np.float64)
                    m = np.asarray([self.mean(x) for x in X], dtype=np.float64)
                    h = c.data_create(X, y, m)
                    self._attach_basis_on(c, h, X)
                    self._peers[i] = (c, h)
            ctxs = [self.ctx] + [c for c, _ in self._peers]
            datas = [self._device_data()] + [h for _, h in self._peers]
            e, status = _capi.Context.energy_batch_multi(ctxs, datas, specs)
        else:
            e, status = self.ctx.energy_batch(self._device_data(), specs)
        bad = [int(s) for s in status if s not in (_capi.DGP_OK, _capi.DGP_ERR_NOT_PD)]
        if bad:
            raise _capi.DgpError(bad[0], "energy batch: a candidate could not be evaluated")
        return [float(v) for v in e]

    def gradEnergy(self, h: Dict[str, float]) -> Dict[str, float]:
        """tr((alpha alpha' - K^-1) dK/dtheta) per effective hyper-parameter, keys un-prefixed (:196-267)."""
        self.covariance.setHyperParameters(h)
        self.noiseModel.setHyperParameters(h)
        _, g = self.ctx.energy_grad(self._device_data(), self._spec())
        g = np.array(g, dtype=np.float64)
        # kernels that re-parametrise a device kind translate their own derivatives (e.g. CoRegRBFKernel)
        off = 0
        for kern in [self.covariance] + ([] if self._dirac_noise else [self.noiseModel]):
            for leaf, cfg in zip(kern._leaves(), kern._leaf_configs(kern.state)):
                nh = len(leaf.hyper_parameters)
                g[off:off + nh] = leaf._fix_device_grad(cfg, [float(v) for v in g[off:off + nh]])
                off += nh
        cov_names = list(self.covariance.hyper_parameters)
        # hParams = (covariance + noiseModel).effective_hyper_parameters, keys stripped of the
        # "<Class@hash>/" prefix (:239); a later key overwrites an earlier one, like Scala's toMap
        pairs = [(h, float(g[cov_names.index(h)])) for h in self.covariance.effective_hyper_parameters]
        if self._dirac_noise:
            pairs += [(h, float(g[len(cov_names)])) for h in self.noiseModel.effective_hyper_parameters]
        else:
            noise_names = list(self.noiseModel.hyper_parameters)
            pairs += [(h, float(g[len(cov_names) + noise_names.index(h)]))
                      for h in self.noiseModel.effective_hyper_parameters]
        return dict(pairs)

    # ---- persistence (:399-422) -------------------------------------------------------------
    def persist(self, state: Dict[str, float]) -> None:
        self.setState(state)
        self._drop_model()
        self._model = self._fit(self._device_data_train_only())
        self.caching = True

    def unpersist(self) -> None:
        self._drop_model()
        self.caching = False

    def _device_data_train_only(self):
        if not self.validationSet:
            return self._device_data()
        X = _as_points([p[0] for p in self.g])
        y = np.asarray([p[1] for p in self.g], dtype=np.float64)
        m = np.asarray([self.mean(x) for x in X], dtype=np.float64)
        if self._train_data is None:
            self._train_data = self.ctx.data_create(X, y, m)
            self._attach_basis(self._train_data, X)
        return self._train_data

    # ---- prediction (:304-393) --------------------------------------------------------------
    def _fitted(self):
        if self._model is not None and self.caching:
            return self._model, False
        return self._fit(self._device_data_train_only()), True

    def predictiveDistribution(self, test: Sequence) -> MultGaussianPRV:
        Xs = _as_points(test)
        ms = np.asarray([self.mean(x) for x in Xs], dtype=np.float64)
        model, temp = self._fitted()
        try:
            self._attach_test_basis(model, Xs)
            mean, cov, _, _ = self.ctx.predict(model, Xs, ms, want_cov=True, want_bars=False)
        finally:
            if temp:
                self.ctx.model_destroy(model)

        def bars(s: float):
            # evaluated lazily, like `root` in BlockedMultiVariateGaussian (:32)
            m2, t2 = self._fitted()
            try:
                self._attach_test_basis(m2, Xs)
                _, _, lo, hi = self.ctx.predict(m2, Xs, ms, want_cov=False, want_bars=True, sigma=s)
            finally:
                if t2:
                    self.ctx.model_destroy(m2)
            return lo, hi

        return MultGaussianPRV(mean, cov, bars)

    def predictionWithErrorBars(self, testData: Sequence, sigma: int):
        Xs = _as_points(testData)
        ms = np.asarray([self.mean(x) for x in Xs], dtype=np.float64)
        model, temp = self._fitted()
        try:
            self._attach_test_basis(model, Xs)
            mean, _, lo, hi = self.ctx.predict(model, Xs, ms, want_cov=False, want_bars=True, sigma=float(sigma))
        finally:
            if temp:
                self.ctx.model_destroy(model)
        return [(testData[i], float(mean[i]), float(lo[i]), float(hi[i])) for i in range(len(testData))]

    def predict(self, point) -> float:
        return self.predictionWithErrorBars([point], 1)[0][1]

    def test(self, testData: Sequence[Tuple[Sequence[float], float]]):
        """ContinuousProcessModel.test (CORE/models/StochasticProcessModel.scala:135-146)."""
        preds = self.predictionWithErrorBars([p[0] for p in testData], 1)
        return [(x, float(t[1]), m, lo, hi) for (x, m, lo, hi), t in zip(preds, testData)]


GPRegression = CudaGPRegression


class GPBasisFuncRegression(CudaGPRegression):
    """GPBasisFuncRegressionModel (CORE/models/gp/GPBasisFuncRegressionModel.scala:70-109): explicit basis functions
    parameterise the trend, with a Gaussian prior N(b, covB) on their coefficients.

    mean(x) = basisFunc(x).b, and the training / cross / test kernel matrices gain
    feature_map_cov(x, y) = (lowB basisFunc(x)).(lowB basisFunc(y)), lowB = cholesky(covB) -- written as in the
    reference (which therefore adds phi' lowB' lowB phi).  `basisFunc` is an arbitrary host function, so the features
    are evaluated here and handed to the device, which applies them as a rank-m tensor-core update of K.
    gradEnergy is inherited from AbstractGPRegressionModel, which differentiates covariance + noiseModel *without*
    the feature-map term (:196-253); grad_with_basis=True uses the full training matrix instead."""

    def __init__(self, cov, noise, trainingdata, basisFunc: Callable[[np.ndarray], Sequence[float]],
                 basis_param_prior: Tuple[Sequence[float], Sequence[Sequence[float]]], ctx=None,
                 grad_with_basis: bool = False):
        b, covB = basis_param_prior
        self.b = np.asarray(b, dtype=np.float64).ravel()
        self.covB = np.asarray(covB, dtype=np.float64)
        if self.covB.shape != (self.b.size, self.b.size):
            raise ValueError("basis_param_prior: covariance must be m x m for m trend coefficients")
        self.lowB = np.linalg.cholesky(self.covB)
        self.basisFunc = basisFunc
        self.grad_with_basis = grad_with_basis
        self._plain = None
        super().__init__(cov, noise, trainingdata,
                         meanFunc=lambda x: float(np.dot(np.asarray(basisFunc(x), dtype=np.float64), self.b)), ctx=ctx)

    def _features(self, X: np.ndarray) -> np.ndarray:
        phi = np.asarray([np.asarray(self.basisFunc(x), dtype=np.float64).ravel() for x in X])
        if phi.ndim != 2 or phi.shape[1] != self.b.size:
            raise ValueError("basisFunc must return as many features as there are trend coefficients")
        return np.ascontiguousarray(phi @ self.lowB.T)  # row i = (lowB phi(x_i))'

    def _attach_basis(self, handle, X):
        self.ctx.data_set_basis(handle, self._features(X))

    def _attach_basis_on(self, ctx, handle, X):
        ctx.data_set_basis(handle, self._features(X))

    def _attach_test_basis(self, model, Xs):
        self.ctx.model_set_test_basis(model, self._features(Xs))

    def gradEnergy(self, h):
        if self.grad_with_basis:
            return super().gradEnergy(h)
        # reference behaviour: a second resident copy of the data without the feature-map term
        if self._plain is None:
            pts = self.g + self.validationSet
            X = _as_points([p[0] for p in pts])
            y = np.asarray([p[1] for p in pts], dtype=np.float64)
            m = np.asarray([self.mean(x) for x in X], dtype=np.float64)
            self._plain = self.ctx.data_create(X, y, m)
        full, self._data = self._data, self._plain
        try:
            return super().gradEnergy(h)
        finally:
            self._data = full

    def _drop_data(self):
        super()._drop_data()
        if getattr(self, "_plain", None) is not None:
            self.ctx.data_destroy(self._plain)
            self._plain = None


def AbstractGPRegressionModel(cov, noise, meanFunc=None):
    """AbstractGPRegressionModel(cov, noise, mean)(data, num) factory (:537-553)."""
    def build(data, num):
        assert len(data) == num
        return CudaGPRegression(cov, noise, data, meanFunc)
    return build


LOG_2PI = math.log(2.0 * math.pi)
