"""Multi-output GP regression (CORE/models/gp/MOGPRegressionModel.scala:19-81).

MOGPRegressionModel is AbstractGPRegressionModel over the product index set (I, Int): every training pattern x with
targets y_0 .. y_{q-1} becomes q points ((x, o), y_o), pattern-major (dataAsSeq, :37-43).  Here the output index is
the last coordinate of the input vector and the kernels are kernels on the concatenated vector, typically
KroneckerProductKernel(k_x, coReg, split = d); the noise model is any kernel (general noise path).

KroneckerMOGPModel exploits the Kronecker structure in `energy` as the reference does (:46-80): the negative log pdf
of MatrixNormal(M, U = K_cov + K_noise over the n inputs, V = coRegCov over the q outputs), which needs one n x n
Cholesky and q solves (dgp_fit, dgp_model_solve) instead of an nq x nq factorisation."""
from __future__ import annotations

import math
from typing import Callable, Dict, Sequence, Tuple

import numpy as np

from . import _capi
from .kernels import DiracKernel, KroneckerProductKernel, LocalScalarKernel, _as_points
from .models import CudaGPRegression


class MOGPRegressionModel(CudaGPRegression):
    def __init__(self, cov: LocalScalarKernel, noise: LocalScalarKernel,
                 data: Sequence[Tuple[Sequence[float], Sequence[float]]], numOutputs: int,
                 meanFunc: Callable[[Tuple[np.ndarray, int]], float] | None = None, ctx=None):
        self.noutputs = int(numOutputs)
        self.patterns = [(np.asarray(x, dtype=np.float64).ravel(), np.asarray(y, dtype=np.float64).ravel())
                         for x, y in data]
        if any(y.size != self.noutputs for _, y in self.patterns):
            raise ValueError("every pattern needs numOutputs targets")
        flat = [(np.append(x, float(o)), float(y[o])) for x, y in self.patterns for o in range(self.noutputs)]
        mf = None if meanFunc is None else (lambda z: float(meanFunc((z[:-1], int(z[-1])))))
        super().__init__(cov, noise, flat, mf, ctx)

    @staticmethod
    def index(x: Sequence[float], output: int) -> np.ndarray:
        """The (x, output) member of the index set as the model's input vector."""
        return np.append(np.asarray(x, dtype=np.float64).ravel(), float(output))


class KroneckerMOGPModel(MOGPRegressionModel):
    """MOGPRegressionModel(covFunc :* coRegCov, noiseCovFunc :* coRegCov, ...) with the matrix-normal energy."""

    def __init__(self, covFunc: LocalScalarKernel, noiseCovFunc: DiracKernel, coRegCov: LocalScalarKernel,
                 data: Sequence[Tuple[Sequence[float], Sequence[float]]], numOutputs: int,
                 meanFunc: Callable[[Tuple[np.ndarray, int]], float] | None = None, ctx=None):
        if not isinstance(noiseCovFunc, DiracKernel):
            raise NotImplementedError("the matrix-normal energy is built for a DiracKernel noise over the inputs")
        d = len(np.asarray(data[0][0], dtype=np.float64).ravel())
        self.covFunc, self.noiseCovFunc, self.coRegCov = covFunc, noiseCovFunc, coRegCov
        self._mean_pair = meanFunc
        super().__init__(KroneckerProductKernel(covFunc, coRegCov, d), KroneckerProductKernel(noiseCovFunc, coRegCov, d),
                         data, numOutputs, meanFunc, ctx)
        self._features = _as_points([x for x, _ in self.patterns])
        self._targets = np.array([y for _, y in self.patterns])
        self._input_data = None

    def __del__(self):
        try:
            if self._input_data is not None:
                self.ctx.data_destroy(self._input_data)
        except Exception:
            pass
        super().__del__()

    def energy(self, h: Dict[str, float], options: Dict[str, str] | None = None) -> float:
        """-MatrixNormal(meanMat, covMatrix + noiseMatrix, colCovMatrix).logPdf(targets)  (:57-79;
        CORE/probability/distributions/MatrixNormal.scala:49-62): 0.5 tr(V^-1 D' U^-1 D)
        + 0.5 (n q log 2 pi + q sum log diag chol U + n sum log diag chol V)."""
        self.setState(h)
        n, q = self._features.shape[0], self.noutputs
        if self._input_data is None:
            self._input_data = self.ctx.data_create(self._features, np.zeros(n))
        mean_mat = np.zeros((n, q)) if self._mean_pair is None else \
            np.array([[self._mean_pair((x, o)) for o in range(q)] for x in self._features])
        D = self._targets - mean_mat
        V = self.ctx.kernel_matrix(self.coRegCov._spec(), np.arange(q, dtype=np.float64)[:, None])
        try:
            rootv = np.linalg.cholesky(V)
            model = self.ctx.fit(self._input_data, self.covFunc._spec(noise=self.noiseCovFunc.state["noiseLevel"]))
        except (np.linalg.LinAlgError, _capi.NotPositiveDefinite):
            return float("inf")
        try:
            _, det_u = self.ctx.model_terms(model)
            Y = self.ctx.model_solve(model, D)
        finally:
            self.ctx.model_destroy(model)
        det_v = float(np.log(np.diag(rootv)).sum())
        inner = D.T @ Y
        tr = float(np.trace(np.linalg.solve(rootv.T, np.linalg.solve(rootv, inner))))
        return 0.5 * tr + 0.5 * (math.log(2.0 * math.pi) * n * q + det_u * q + det_v * n)
