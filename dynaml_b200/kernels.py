"""Host-side mirror of dynaml-core's covariance-function API for the GP hot path.

Same names, argument meaning and error behaviour as the reference classes
(CORE = dynaml-core/src/main/scala/io/github/tailhq/dynaml):

  CovarianceFunction / LocalScalarKernel   CORE/kernels/CovarianceFunction.scala:17-59,
                                           CORE/kernels/LocalScalarKernel.scala:35-176
  SEKernel, RBFKernel, MahalanobisKernel   CORE/kernels/RBFKernel.scala:29-146
  GenericMaternKernel                      CORE/kernels/GenericMaternKernel.scala:37-99
  PeriodicKernel                           CORE/kernels/PeriodicKernel.scala:11-68
  DiracKernel                              CORE/kernels/DiracKernel.scala:30-60
  AdditiveCovariance                       CORE/kernels/AdditiveCovariance.scala:27-93

The classes only carry hyper-parameter state; every number is produced by the CUDA library
through the C ABI (dynaml_b200._capi).  Kernels the device path does not implement raise
NotImplementedError -- there is no CPU fallback.  CauchyKernel, LaplacianKernel, RationalQuadraticKernel
and TStudentKernel ride the same builders as extra assembly epilogues.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence

import numpy as np

from . import _capi


def _as_points(data) -> np.ndarray:
    """Seq[DenseVector[Double]] -> N x d array."""
    a = np.asarray(data, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError("expected a sequence of d-vectors")
    return np.ascontiguousarray(a)


class SVMKernelMatrix:
    """CORE/kernels/SVMKernelMatrix.scala -- holder returned by buildKernelMatrix."""

    def __init__(self, kernel: np.ndarray, dimension: int):
        self.kernel = kernel
        self.rows = self.cols = dimension

    def getKernelMatrix(self) -> np.ndarray:
        return self.kernel


class LocalScalarKernel:
    """CovarianceFunction[DenseVector[Double], Double, DenseMatrix[Double]] with LocalScalarKernel."""

    hyper_parameters: List[str] = []

    def __init__(self):
        self.blocked_hyper_parameters: List[str] = []
        self.state: Dict[str, float] = {}
        self.rowBlocking, self.colBlocking = 1000, 1000
        self.grad_mode = _capi.GRAD_REFERENCE

    # ---- hyper-parameter plumbing (CovarianceFunction.scala:21-44) ------------------------
    def block(self, *h: str) -> None:
        self.blocked_hyper_parameters = list(h)

    def block_all_hyper_parameters(self) -> None:
        self.blocked_hyper_parameters = list(self.hyper_parameters)

    @property
    def effective_state(self) -> Dict[str, float]:
        return {k: v for k, v in self.state.items() if k not in self.blocked_hyper_parameters}

    @property
    def effective_hyper_parameters(self) -> List[str]:
        return [h for h in self.hyper_parameters if h not in self.blocked_hyper_parameters]

    def setHyperParameters(self, h: Dict[str, float]):
        assert all(k in h for k in self.effective_hyper_parameters), \
            "All hyper parameters must be contained in the arguments"
        for key in self.effective_hyper_parameters:
            self.state[key] = float(h[key])
        return self

    def setBlockSizes(self, s) -> None:
        self.rowBlocking, self.colBlocking = int(s[0]), int(s[1])

    def __str__(self) -> str:  # the reference's Object.toString: "<package>.<Class>@<hash>"
        return f"io.github.tailhq.dynaml.kernels.{type(self).__name__}@{id(self) & 0xffffffff:x}"

    # ---- device description ---------------------------------------------------------------
    def _device_kind(self) -> int:
        raise NotImplementedError(
            f"{type(self).__name__} has no sm_100a implementation (supported: SEKernel, RBFKernel, "
            "MahalanobisKernel, GenericMaternKernel, PeriodicKernel, DiracKernel, CauchyKernel, LaplacianKernel, "
            "RationalQuadraticKernel, TStudentKernel, PolynomialKernel, FBMKernel and their sums / products); "
            "no CPU fallback")

    def _device_hyp(self, config: Dict[str, float]) -> List[float]:
        return [float(config[h]) for h in self.hyper_parameters]

    def _spec(self, config: Dict[str, float] | None = None, noise: float = 0.0):
        cfg = self.state if config is None else config
        return _capi.make_spec(self._device_kind(), self._device_hyp(cfg), noise, self.grad_mode)

    # ---- evaluation -----------------------------------------------------------------------
    def evaluateAt(self, config: Dict[str, float]):
        def f(x, y) -> float:
            K = _capi.default_context().kernel_matrix(self._spec(config), _as_points([x]), _as_points([y]))
            return float(K[0, 0])
        return f

    def evaluate(self, x, y) -> float:
        return self.evaluateAt(self.state)(x, y)

    def buildKernelMatrix(self, mappedData: Sequence, length: int) -> SVMKernelMatrix:
        X = _as_points(mappedData)
        assert X.shape[0] == length
        return SVMKernelMatrix(_capi.default_context().kernel_matrix(self._spec(), X), length)

    def buildCrossKernelMatrix(self, dataset1: Sequence, dataset2: Sequence) -> np.ndarray:
        return _capi.default_context().kernel_matrix(self._spec(), _as_points(dataset1), _as_points(dataset2))

    def buildBlockedKernelMatrix(self, mappedData: Sequence, length: int):
        """LocalScalarKernel.scala:160-161: the PartitionedPSDMatrix of blocks rowBlocking x colBlocking."""
        return SVMKernel.buildPartitionedKernelMatrix(mappedData, length, self.rowBlocking, self.colBlocking, self)

    def buildBlockedCrossKernelMatrix(self, dataset1: Sequence, dataset2: Sequence):
        """LocalScalarKernel.scala:163-164."""
        return SVMKernel.crossPartitonedKernelMatrix(dataset1, dataset2, self.rowBlocking, self.colBlocking, self)

    # ---- kernel algebra (LocalScalarKernel.scala:53-91) -------------------------------------
    # Every kernel reduces to the normal form  sum_t coef_t * prod_f leaf_f ^ e_tf  that the device
    # evaluates (DGP_KERNEL_COMPOSITE): `_leaves()` lists the base kernels in hyper-parameter order,
    # `_terms(config)` the terms as (coef, {leaf position: exponent}).
    def _leaves(self) -> List["LocalScalarKernel"]:
        return [self]

    def _leaf_configs(self, config: Dict[str, float]) -> List[Dict[str, float]]:
        return [config]

    def _terms(self) -> List[tuple]:
        return [(1.0, {0: 1})]

    def _leaf_ranges(self) -> List[tuple]:
        """Per leaf the (start, count) slice of the input vector it sees; count 0 = up to the last coordinate."""
        return [(0, 0)]

    def _fix_device_grad(self, config: Dict[str, float], g: List[float]) -> List[float]:
        """Map the device derivatives (w.r.t. `_device_hyp`) to the reference's for this kernel's own
        hyper-parameters; identity unless the kernel is a re-parametrisation of a device kind."""
        return g

    def __add__(self, other: "LocalScalarKernel") -> "AdditiveCovariance":
        return AdditiveCovariance(self, other)

    def __mul__(self, other):
        if isinstance(other, LocalScalarKernel):
            return MultiplicativeCovariance(self, other)
        return ScaledKernel(self, float(other))

    __rmul__ = __mul__


class SVMKernel:
    """The matrix builders of the SVMKernel companion object (CORE/kernels/SVMKernel.scala:71-307).  The reference
    passes per-pair closures `eval` / `evalGrad`; here the kernel object itself is passed and the device evaluates
    it (`noise`, optional, is the level of a DiracKernel noise model added on coincident points, i.e. the
    `cov + noise` eval the GP model hands over; AbstractGPRegressionModel.scala:269-277)."""

    @staticmethod
    def buildSVMKernelMatrix(mappedData: Sequence, length: int, kernel: "LocalScalarKernel", noise: float | None = None):
        X = _as_points(mappedData)
        assert X.shape[0] == length
        K = _capi.default_context().kernel_matrix(kernel._spec(noise=0.0 if noise is None else noise), X,
                                                  add_noise=noise is not None)
        return SVMKernelMatrix(K, length)

    @staticmethod
    def crossKernelMatrix(data1: Sequence, data2: Sequence, kernel: "LocalScalarKernel") -> np.ndarray:
        return _capi.default_context().kernel_matrix(kernel._spec(), _as_points(data1), _as_points(data2))

    @staticmethod
    def buildPartitionedKernelMatrix(data: Sequence, length: int, numElementsPerRowBlock: int, numElementsPerColBlock: int,
                                     kernel: "LocalScalarKernel", noise: float | None = None, ctx=None):
        """:181-219.  Lower block list (block row >= block column) of the training matrix, every block cut straight
        from the points resident on the device (dgp_data_kernel_block: no re-upload per block).  Diagonal blocks
        require equal row and column block sizes, as every caller of the reference uses them."""
        from .algebra import PartitionedPSDMatrix
        ctx = ctx or _capi.default_context()
        X = _as_points(data)
        assert X.shape[0] == length
        nr, nc = int(numElementsPerRowBlock), int(numElementsPerColBlock)
        if nr != nc:
            raise ValueError("buildPartitionedKernelMatrix: equal row and column block sizes are required "
                             "(a PartitionedPSDMatrix has square diagonal blocks)")
        spec = kernel._spec(noise=0.0 if noise is None else noise)
        h = ctx.data_create(X, np.zeros(length))
        try:
            nb = -(-length // nr)
            blocks = []
            for i in range(nb):
                for j in range(i + 1):
                    r0, c0 = i * nr, j * nc
                    blocks.append(((i, j), ctx.data_kernel_block(h, spec, r0, min(nr, length - r0), c0, min(nc, length - c0),
                                                                 add_noise=noise is not None)))
        finally:
            ctx.data_destroy(h)
        return PartitionedPSDMatrix(blocks, length, length, nb, nb)

    @staticmethod
    def crossPartitonedKernelMatrix(data1: Sequence, data2: Sequence, numElementsPerRowBlock: int,
                                    numElementsPerColBlock: int, kernel: "LocalScalarKernel", ctx=None):
        """:221-248 (the reference's spelling).  Every block is a crossKernelMatrix of the two groups."""
        from .algebra import PartitionedMatrix
        ctx = ctx or _capi.default_context()
        X1, X2 = _as_points(data1), _as_points(data2)
        nr, nc = int(numElementsPerRowBlock), int(numElementsPerColBlock)
        rb, cb = -(-X1.shape[0] // nr), -(-X2.shape[0] // nc)
        spec = kernel._spec()
        blocks = [((i, j), ctx.kernel_matrix(spec, X1[i * nr:(i + 1) * nr], X2[j * nc:(j + 1) * nc]))
                  for i in range(rb) for j in range(cb)]
        return PartitionedMatrix(blocks, X1.shape[0], X2.shape[0], rb, cb)

    @staticmethod
    def _grad_names(kernel: "LocalScalarKernel", noise: float | None) -> List[str]:
        names = []
        for leaf in kernel._leaves():
            names += list(leaf.hyper_parameters)
        return names + ["noiseLevel"]

    @staticmethod
    def buildKernelGradMatrix(data1: Sequence, hyper_parameters: Sequence[str], kernel: "LocalScalarKernel",
                              noise: float | None = None, ctx=None) -> Dict[str, np.ndarray]:
        """:107-145: {"kernel-matrix": K, h: dK/dh for h in hyper_parameters} (h also "noiseLevel" with a noise model)."""
        ctx = ctx or _capi.default_context()
        X = _as_points(data1)
        K, G = ctx.kernel_grad_matrix(kernel._spec(noise=0.0 if noise is None else noise), X, add_noise=noise is not None)
        return SVMKernel._as_map(kernel, noise, K, G, hyper_parameters)

    @staticmethod
    def buildCrossKernelGradMatrix(data1: Sequence, data2: Sequence, hyper_parameters: Sequence[str],
                                   kernel: "LocalScalarKernel", ctx=None) -> Dict[str, np.ndarray]:
        """:147-179."""
        ctx = ctx or _capi.default_context()
        K, G = ctx.kernel_grad_matrix(kernel._spec(), _as_points(data1), _as_points(data2))
        return SVMKernel._as_map(kernel, None, K, G, hyper_parameters)

    @staticmethod
    def _as_map(kernel, noise, K, G, hyper_parameters) -> Dict[str, np.ndarray]:
        out = {"kernel-matrix": K}
        # device order: the hyper-parameters of the kernel's leaves, then the noise level; kernels that re-parametrise a
        # device kind translate their own derivatives (as gradEnergy does)
        off = 0
        by_name: Dict[str, np.ndarray] = {}
        for leaf, cfg in zip(kernel._leaves(), kernel._leaf_configs(kernel.state)):
            nh = len(leaf.hyper_parameters)
            planes = G[off:off + nh]
            if type(leaf)._fix_device_grad is not LocalScalarKernel._fix_device_grad:
                stack = np.stack(planes)
                fixed = np.apply_along_axis(lambda v: np.asarray(leaf._fix_device_grad(cfg, list(v))), 0, stack)
                planes = [fixed[q] for q in range(nh)]
            for name, pl in zip(leaf.hyper_parameters, planes):
                by_name[name] = pl
            off += nh
        if noise is not None:
            by_name["noiseLevel"] = G[off]
        for h in hyper_parameters:
            key = h.split("/")[-1]
            if key not in by_name:
                raise KeyError(f"no derivative for hyper-parameter {h!r}")
            out[h] = by_name[key]
        return out


def composite_spec(parts: Sequence[tuple], noise: float = 0.0, extra_terms: Sequence[tuple] = ()):
    """DGP_KERNEL_COMPOSITE spec of a sum of kernels: `parts` = [(kernel, config), ...]; `extra_terms` adds
    (coefficient, {leaf position: exponent}) terms over the joint leaf list (e.g. a constant: (c, {}))."""
    factors, terms, ranges, off = [], [], [], 0
    for kernel, config in parts:
        leaves, cfgs = kernel._leaves(), kernel._leaf_configs(config)
        factors += [(k._device_kind(), k.grad_mode, k._device_hyp(c)) for k, c in zip(leaves, cfgs)]
        terms += [(c, {f + off: e for f, e in m.items()}) for c, m in kernel._terms()]
        ranges += list(kernel._leaf_ranges())
        off += len(leaves)
    terms += list(extra_terms)
    dense = [(c, [m.get(f, 0) for f in range(off)]) for c, m in terms]
    return _capi.make_composite_spec(factors, dense, noise, ranges if any(r != (0, 0) for r in ranges) else None)


class RBFKernel(LocalScalarKernel):
    hyper_parameters = ["bandwidth"]

    def __init__(self, bandwidth: float = 1.0):
        super().__init__()
        self.state = {"bandwidth": float(bandwidth)}

    def setbandwidth(self, d: float) -> None:
        self.state["bandwidth"] = float(d)

    def _device_kind(self):
        return _capi.KERNEL_RBF


class SEKernel(RBFKernel):
    hyper_parameters = ["bandwidth", "amplitude"]

    def __init__(self, band: float = 1.0, h: float = 2.0):
        super().__init__(band)
        self.state = {"bandwidth": float(band), "amplitude": float(h)}

    def _device_kind(self):
        return _capi.KERNEL_SE


class MahalanobisKernel(LocalScalarKernel):
    """ARD squared-exponential.  grad_mode selects the reference's (defective) or the
    exact bandwidth derivative; the default reproduces the reference."""

    def __init__(self, band: Iterable[float], h: float = 2.0, grad_mode: int = _capi.GRAD_REFERENCE):
        super().__init__()
        band = [float(b) for b in band]
        self.hyper_parameters = ["MahalanobisAmplitude"] + [f"MahalanobisBandwidth_{i}" for i in range(len(band))]
        self.state = {"MahalanobisAmplitude": float(h)}
        self.state.update({f"MahalanobisBandwidth_{i}": b for i, b in enumerate(band)})
        self.grad_mode = grad_mode

    def _device_kind(self):
        return _capi.KERNEL_ARD_SE


class GenericMaternKernel(LocalScalarKernel):
    hyper_parameters = ["p", "l"]

    def __init__(self, l: float, p: int = 2):
        super().__init__()
        self.blocked_hyper_parameters = ["p"]
        self.state = {"p": float(p), "l": float(l)}

    def setHyperParameters(self, h: Dict[str, float]):
        super().setHyperParameters(h)
        if "p" in h:
            self.state["p"] = float(math.floor(abs(h["p"])))
        return self

    def _device_kind(self):
        return _capi.KERNEL_MATERN


class PeriodicKernel(LocalScalarKernel):
    hyper_parameters = ["sigma", "lengthscale", "frequency"]

    def __init__(self, s: float = 1.0, ls: float = 1.0, freq: float = 1.0):
        super().__init__()
        self.state = {"sigma": float(s), "lengthscale": float(ls), "frequency": float(freq)}

    def setlengthscale(self, d: float) -> None:
        self.state["lengthscale"] = float(d)

    def setfrequency(self, f: float) -> None:
        self.state["frequency"] = float(f)

    def _device_kind(self):
        return _capi.KERNEL_PERIODIC


class CauchyKernel(LocalScalarKernel):
    """CORE/kernels/CauchyKernel.scala: 1 / (1 + (|x-y|_2 / sigma)^2)."""
    hyper_parameters = ["sigma"]

    def __init__(self, si: float = 1.0):
        super().__init__()
        self.state = {"sigma": float(si)}

    def setsigma(self, b: float) -> None:
        self.state["sigma"] = float(b)

    def _device_kind(self):
        return _capi.KERNEL_CAUCHY


class LaplacianKernel(LocalScalarKernel):
    """CORE/kernels/LaplacianKernel.scala: exp(-|x-y|_1 / beta)."""
    hyper_parameters = ["beta"]

    def __init__(self, be: float = 1.0):
        super().__init__()
        self.state = {"beta": float(be)}

    def setbeta(self, b: float) -> None:
        self.state["beta"] = float(b)

    def _device_kind(self):
        return _capi.KERNEL_LAPLACIAN


class RationalQuadraticKernel(LocalScalarKernel):
    """CORE/kernels/RationalQuadraticKernel.scala: (1 + |x-y|^2 / (mu l^2))^(-(d + mu)/2)."""
    hyper_parameters = ["mu", "l"]

    def __init__(self, shape: float = 1.0, l: float = 1.0):
        super().__init__()
        self.state = {"mu": float(shape), "l": float(l)}

    def setShape(self, b: float) -> None:
        self.state["mu"] = float(b)

    def setScale(self, lam: float) -> None:
        self.state["l"] = float(lam)

    def _device_kind(self):
        return _capi.KERNEL_RATQUAD


class TStudentKernel(LocalScalarKernel):
    """CORE/kernels/TStudentKernel.scala: 1 / (1 + |x-y|_2^d)."""
    hyper_parameters = ["d"]

    def __init__(self, degree: float = 1.0):
        super().__init__()
        self.state = {"d": float(degree)}

    @property
    def d(self) -> float:
        return self.state["d"]

    def _device_kind(self):
        return _capi.KERNEL_TSTUDENT


class PolynomialKernel(LocalScalarKernel):
    """k = (x.y + offset)^int(degree) (PolynomialKernel.scala:9-45); setHyperParameters stores |offset|; the
    reference defines no gradient for it (stub zeros)."""
    hyper_parameters = ["degree", "offset"]

    def __init__(self, degree: int = 2, offset: float = 1.0):
        super().__init__()
        self.state = {"degree": float(degree), "offset": float(offset)}

    def setdegree(self, d: int) -> None:
        self.state["degree"] = float(d)

    def setoffset(self, o: float) -> None:
        self.state["offset"] = float(o)

    def setHyperParameters(self, h: Dict[str, float]):
        super().setHyperParameters(h)
        if "offset" in h:
            self.state["offset"] = abs(float(h["offset"]))
        return self

    def _device_kind(self):
        return _capi.KERNEL_POLYNOMIAL


class FBMKernel(LocalScalarKernel):
    """Fractional Brownian motion covariance (|x|^2H + |y|^2H - |x-y|^2H)/2 (FBMKernel.scala:70-122)."""
    hyper_parameters = ["hurst"]

    def __init__(self, hurst: float = 0.75, grad_mode: int = _capi.GRAD_REFERENCE):
        super().__init__()
        self.state = {"hurst": float(hurst)}
        self.grad_mode = grad_mode

    def sethurst(self, d: float) -> None:
        self.state["hurst"] = float(d)

    def getHurst(self) -> float:
        return self.state["hurst"]

    def _device_kind(self):
        return _capi.KERNEL_FBM


class DiracKernel(LocalScalarKernel):
    hyper_parameters = ["noiseLevel"]

    def __init__(self, noiseLevel: float = 1.0):
        super().__init__()
        self.state = {"noiseLevel": float(noiseLevel)}

    def setNoiseLevel(self, d: float) -> None:
        self.state["noiseLevel"] = float(d)

    def _device_kind(self):
        return _capi.KERNEL_DIRAC

    def buildKernelMatrix(self, mappedData: Sequence, length: int) -> SVMKernelMatrix:
        # DiracKernel.scala:55-58 overrides the builder: eye(length) * noiseLevel (signed, no abs)
        return SVMKernelMatrix(np.eye(length) * self.state["noiseLevel"], length)


class CompositeCovariance(LocalScalarKernel):
    """Shared plumbing of AdditiveCovariance / MultiplicativeCovariance: "<Class@hash>/<hyp>" naming and
    substring routing of configurations (AdditiveCovariance.scala:32-58, CompositeCovariance.truncateState),
    plus the reduction to the device's sum-of-products form."""

    def __init__(self, firstKernel: LocalScalarKernel, otherKernel: LocalScalarKernel):
        super().__init__()
        self.firstKernel, self.otherKernel = firstKernel, otherKernel
        self.fID = str(firstKernel).split(".")[-1]
        self.sID = str(otherKernel).split(".")[-1]
        self.hyper_parameters = [f"{self.fID}/{h}" for h in firstKernel.hyper_parameters] + \
                                [f"{self.sID}/{h}" for h in otherKernel.hyper_parameters]
        self.blocked_hyper_parameters = [f"{self.fID}/{h}" for h in firstKernel.blocked_hyper_parameters] + \
                                        [f"{self.sID}/{h}" for h in otherKernel.blocked_hyper_parameters]
        self.state = {f"{self.fID}/{k}": v for k, v in firstKernel.state.items()}
        self.state.update({f"{self.sID}/{k}": v for k, v in otherKernel.state.items()})

    @staticmethod
    def _truncate(k: str) -> str:
        return "/".join(k.split("/")[1:])

    def _configs(self, config: Dict[str, float]):
        return ({self._truncate(k): v for k, v in config.items() if self.fID in k},
                {self._truncate(k): v for k, v in config.items() if self.sID in k})

    def setHyperParameters(self, h: Dict[str, float]):
        c1, c2 = self._configs(h)
        self.firstKernel.setHyperParameters(c1)
        self.otherKernel.setHyperParameters(c2)
        return super().setHyperParameters(h)

    def block(self, *h: str) -> None:
        self.firstKernel.block(*[self._truncate(k) for k in h if self.fID in k])
        self.otherKernel.block(*[self._truncate(k) for k in h if self.sID in k])
        super().block(*h)

    # ---- normal form ----------------------------------------------------------------------
    def _device_kind(self) -> int:
        for leaf in self._leaves():
            leaf._device_kind()  # raises for base kernels without a device implementation
        return _capi.KERNEL_COMPOSITE

    def _leaves(self):
        return self.firstKernel._leaves() + self.otherKernel._leaves()

    def _leaf_configs(self, config):
        c1, c2 = self._configs(config)
        return self.firstKernel._leaf_configs(c1) + self.otherKernel._leaf_configs(c2)

    def _leaf_ranges(self):
        return self.firstKernel._leaf_ranges() + self.otherKernel._leaf_ranges()

    def _shifted_other_terms(self):
        off = len(self.firstKernel._leaves())
        return [(c, {f + off: e for f, e in m.items()}) for c, m in self.otherKernel._terms()]

    def _spec(self, config=None, noise: float = 0.0):
        return composite_spec([(self, self.state if config is None else config)], noise)


class AdditiveCovariance(CompositeCovariance):
    """k = k1 + k2 (AdditiveCovariance.scala:27-93).  `<kernel> + DiracKernel` keeps the dedicated noise path of
    the single-kernel spec (the GP model's cov + noise); every other sum runs as a device composite."""

    def _terms(self):
        return self.firstKernel._terms() + self._shifted_other_terms()

    def _is_cov_plus_noise(self) -> bool:
        return isinstance(self.otherKernel, DiracKernel) and not isinstance(self.firstKernel, CompositeCovariance) \
            and not isinstance(self.firstKernel, ScaledKernel)

    def _spec(self, config=None, noise: float = 0.0, as_model_covariance: bool = False):
        # Inside a GP model the gradient is indexed by this kernel's own hyper-parameter list followed by the
        # model's noise level, so the sum must stay a composite there (its Dirac member is an ordinary factor).
        if self._is_cov_plus_noise() and noise == 0.0 and not as_model_covariance:
            c1, c2 = self._configs(self.state if config is None else config)
            return self.firstKernel._spec(c1, noise=c2["noiseLevel"])
        return super()._spec(config, noise)

    def buildKernelMatrix(self, mappedData: Sequence, length: int) -> SVMKernelMatrix:
        X = _as_points(mappedData)
        return SVMKernelMatrix(_capi.default_context().kernel_matrix(self._spec(), X, add_noise=True), length)

    def buildCrossKernelMatrix(self, dataset1: Sequence, dataset2: Sequence) -> np.ndarray:
        if self._is_cov_plus_noise():
            # crossKernelMatrix evaluates cov + noise pairwise: the Dirac term fires only on exact duplicates
            # across the two sets, which the dedicated noise path does not add.
            raise NotImplementedError("cross matrices of cov + noise are not on the GP hot path")
        return super().buildCrossKernelMatrix(dataset1, dataset2)


class MultiplicativeCovariance(CompositeCovariance):
    """k = k1 * k2 (MultiplicativeCovariance.scala:11-90): every term of k1 times every term of k2."""

    def _terms(self):
        out = []
        for c1, m1 in self.firstKernel._terms():
            for c2, m2 in self._shifted_other_terms():
                m = dict(m1)
                for f, e in m2.items():
                    m[f] = m.get(f, 0) + e
                out.append((c1 * c2, m))
        return out


class ScaledKernel(LocalScalarKernel):
    """k * c for a positive constant (LocalScalarKernel.scala:68-91): shares the hyper-parameters and state of the
    wrapped kernel; value and gradient are scaled by c."""

    def __init__(self, base: LocalScalarKernel, c: float):
        if not c > 0:
            raise ValueError("Multiplicative constant applied on a kernel must be positive!")
        super().__init__()
        self.base, self.c = base, float(c)
        self.hyper_parameters = list(base.hyper_parameters)
        self.state = base.state                      # shared, as in the reference (`state = self.state`)
        self.blocked_hyper_parameters = list(base.blocked_hyper_parameters)

    def setHyperParameters(self, h: Dict[str, float]):
        self.base.setHyperParameters(h)
        return super().setHyperParameters(h)

    def _device_kind(self) -> int:
        for leaf in self._leaves():
            leaf._device_kind()
        return _capi.KERNEL_COMPOSITE

    def _leaves(self):
        return self.base._leaves()

    def _leaf_configs(self, config):
        return self.base._leaf_configs(config)

    def _terms(self):
        return [(self.c * c, m) for c, m in self.base._terms()]

    def _leaf_ranges(self):
        return self.base._leaf_ranges()

    def _spec(self, config=None, noise: float = 0.0):
        return composite_spec([(self, self.state if config is None else config)], noise)


class DecomposableCovariance(LocalScalarKernel):
    """DecomposableCovariance(kernels*)(replication encoder, WeightedSumReducer(weights)) as ProbGPCommMachine builds it
    (CORE/kernels/DecomposableCovariance.scala:28-120, CORE/optimization/ProbGPCommMachine.scala:92-121):
    k(x, y) = sum_i weights_i * k_i(x, y) over the same inputs, hyper-parameters "<Class@hash>/<name>" per member.
    (The reference pairs weights with kernels in the iteration order of an immutable Map keyed by the ids, i.e. in
    hash order for more than four members; the intended pairing -- position i with weight i -- is used here.)"""

    def __init__(self, kernels: Sequence[LocalScalarKernel], weights: Sequence[float] | None = None):
        super().__init__()
        self.kernels = list(kernels)
        if not self.kernels:
            raise ValueError("DecomposableCovariance needs at least one kernel")
        self.weights = [1.0] * len(self.kernels) if weights is None else [float(w) for w in weights]
        if len(self.weights) != len(self.kernels):
            raise ValueError("one weight per kernel")
        self.ids = [str(k).split(".")[-1] for k in self.kernels]
        self.hyper_parameters = [f"{i}/{h}" for i, k in zip(self.ids, self.kernels) for h in k.hyper_parameters]
        self.blocked_hyper_parameters = [f"{i}/{h}" for i, k in zip(self.ids, self.kernels)
                                         for h in k.blocked_hyper_parameters]
        self.state = {f"{i}/{h}": v for i, k in zip(self.ids, self.kernels) for h, v in k.state.items()}

    def setHyperParameters(self, h: Dict[str, float]):
        assert all(k in h for k in self.effective_hyper_parameters), \
            "All hyper parameters must be contained in the arguments"
        for i, k in zip(self.ids, self.kernels):
            sub = {key.split("/", 1)[1]: v for key, v in h.items() if key.split("/")[0] == i and "/" in key}
            if sub:
                k.setHyperParameters(sub)
        for i, k in zip(self.ids, self.kernels):
            for name, v in k.state.items():
                self.state[f"{i}/{name}"] = v
        return self

    def _device_kind(self) -> int:
        for leaf in self._leaves():
            leaf._device_kind()
        return _capi.KERNEL_COMPOSITE

    def _leaves(self):
        return [leaf for k in self.kernels for leaf in k._leaves()]

    def _leaf_configs(self, config):
        out = []
        for i, k in zip(self.ids, self.kernels):
            sub = {key.split("/", 1)[1]: v for key, v in config.items() if key.split("/")[0] == i and "/" in key}
            out += k._leaf_configs(sub)
        return out

    def _leaf_ranges(self):
        return [r for k in self.kernels for r in k._leaf_ranges()]

    def _terms(self):
        out, off = [], 0
        for w, k in zip(self.weights, self.kernels):
            out += [(w * c, {f + off: e for f, e in m.items()}) for c, m in k._terms()]
            off += len(k._leaves())
        return out

    def _spec(self, config=None, noise: float = 0.0):
        return composite_spec([(self, self.state if config is None else config)], noise)


class TensorCombinationKernel(CompositeCovariance):
    """Kernels on tuple inputs (x1, x2): k((x1, x2), (y1, y2)) = reduce(k1(x1, y1), k2(x2, y2))
    (CORE/kernels/TensorCombinationKernel.scala:12-97).  Inputs are the concatenated vectors; `split` is the length
    of the first component, so k1 sees coordinates [0, split) and k2 the rest.  `k1 :* k2` is
    KroneckerProductKernel (product reducer), `k1 :+ k2` the sum reducer."""

    def __init__(self, firstK: LocalScalarKernel, secondK: LocalScalarKernel, split: int, reducer: str = "*"):
        if reducer not in ("*", "+"):
            raise ValueError("reducer must be '*' (Reducer.:*:) or '+' (Reducer.:+:)")
        if split <= 0:
            raise ValueError("split must be the positive length of the first component")
        super().__init__(firstK, secondK)
        self.split, self.reducer = int(split), reducer

    def _terms(self):
        if self.reducer == "+":
            return self.firstKernel._terms() + self._shifted_other_terms()
        out = []
        for c1, m1 in self.firstKernel._terms():
            for c2, m2 in self._shifted_other_terms():
                m = dict(m1)
                for f, e in m2.items():
                    m[f] = m.get(f, 0) + e
                out.append((c1 * c2, m))
        return out

    def _leaf_ranges(self):
        first = []
        for s0, c0 in self.firstKernel._leaf_ranges():
            if s0 >= self.split:
                raise ValueError("a kernel of the first component addresses coordinates beyond the split")
            first.append((s0, c0 if c0 else self.split - s0))
        second = [(self.split + s0, c0) for s0, c0 in self.otherKernel._leaf_ranges()]
        return first + second


class KroneckerProductKernel(TensorCombinationKernel):
    def __init__(self, firstK: LocalScalarKernel, secondK: LocalScalarKernel, split: int):
        super().__init__(firstK, secondK, split, "*")


# ---- co-regionalisation kernels over the output index (LocalScalarKernel[Int] in the reference) -----------------
# In the multi-output models the output index travels as the last coordinate of the input vector, so these are the
# device kernels of the same functional form restricted to that coordinate, with the reference's names.
class CoRegRBFKernel(LocalScalarKernel):
    """exp(-(i - j)^2 / coRegB) (RBFKernel.scala:182-194).  Device: RBF with l = sqrt(coRegB / 2); the reference's
    derivative is k (i-j)^2 / |coRegB|^3 (not the chain rule), i.e. the device's k (i-j)^2 / l^3 times l^3 / |b|^3."""
    hyper_parameters = ["coRegB"]

    def __init__(self, bandwidth: float):
        super().__init__()
        self.state = {"coRegB": float(bandwidth)}

    def _device_kind(self):
        return _capi.KERNEL_RBF

    def _device_hyp(self, config):
        return [math.sqrt(float(config["coRegB"]) / 2.0)]

    def _fix_device_grad(self, config, g):
        b = float(config["coRegB"])
        l = math.sqrt(b / 2.0)
        return [g[0] * l ** 3 / abs(b) ** 3]


class CoRegCauchyKernel(CauchyKernel):
    hyper_parameters = ["CoRegSigma"]        # CauchyKernel.scala:53-65

    def __init__(self, bandwidth: float):
        LocalScalarKernel.__init__(self)
        self.state = {"CoRegSigma": float(bandwidth)}


class CoRegLaplaceKernel(LaplacianKernel):
    hyper_parameters = ["coRegLB"]           # LaplacianKernel.scala:53-65

    def __init__(self, bandwidth: float):
        LocalScalarKernel.__init__(self)
        self.state = {"coRegLB": float(bandwidth)}


class CoRegTStudentKernel(TStudentKernel):
    hyper_parameters = ["CoRegD"]            # TStudentKernel.scala:74-94

    def __init__(self, bandwidth: float):
        LocalScalarKernel.__init__(self)
        self.state = {"CoRegD": float(bandwidth)}


class CoRegDiracKernel(LocalScalarKernel):
    """1 if i == j else 0 (DiracKernel.scala:91-98).  The reference class has no hyper-parameters; the device Dirac
    kind carries a level, kept here as the permanently blocked `coRegDiracLevel` = 1."""
    hyper_parameters = ["coRegDiracLevel"]

    def __init__(self):
        super().__init__()
        self.state = {"coRegDiracLevel": 1.0}
        self.blocked_hyper_parameters = ["coRegDiracLevel"]

    def block(self, *h: str) -> None:
        self.blocked_hyper_parameters = ["coRegDiracLevel"]

    def _device_kind(self):
        return _capi.KERNEL_DIRAC
