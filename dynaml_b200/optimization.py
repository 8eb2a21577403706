"""Host-side mirror of the hyper-parameter tuners that drive the GP hot path.

Reference (CORE/optimization): ModelTuner.scala:40-160, AbstractGridSearch.scala, GridSearch.scala:27-63,
AbstractCSA.scala:30-327, CoupledSimulatedAnnealing.scala:30-54, GradBasedGlobalOptimizer.scala:27-103.

The reference evaluates `system.energy` once per candidate, serially (ModelTuner.scala:131-153,
AbstractCSA.scala:214-259).  Here each landscape / CSA sweep is one `energyBatch` call, and under
torch.distributed (one process per GPU) the candidates of a sweep are sharded round-robin across
ranks and gathered -- the path has no data-path collective.

The reference draws from unseeded global RNGs (scala.util.Random, Breeze CauchyDistribution), so
its CSA trajectories are not reproducible; the RNG is injected here (`rng=`).
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

Config = Dict[str, float]


def _exp(x: float) -> float:
    """math.exp with the JVM's IEEE semantics: overflow gives +Inf instead of raising (scala.math.exp)."""
    with np.errstate(all="ignore"):
        return float(np.exp(np.float64(x)))


def _div(a: float, b: float) -> float:
    """Double division with the JVM's IEEE semantics: x/0 = +-Inf, 0/0 = Inf/Inf = NaN, never an exception."""
    with np.errstate(all="ignore"):
        return float(np.float64(a) / np.float64(b))


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:  # pragma: no cover
        pass
    return None


def broadcast_from_rank0(obj, group=None):
    """Rank 0's copy of a small host object on every rank (identity outside torch.distributed).  Grids drawn from
    priors, CSA proposals and the accept draws come from unseeded RNGs in the reference; under one process per
    GPU they must be ONE rank's draws, or the gathered energies would be paired with other ranks' proposals."""
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return box[0]


def _configs_digest(configs: Sequence[Config]) -> str:
    import hashlib
    h = hashlib.sha256()
    for c in configs:
        for k in sorted(c):
            h.update(k.encode())
            h.update(np.float64(c[k]).tobytes())
        h.update(b"|")
    return h.hexdigest()


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Round-robin candidate -> rank map (candidate i runs on GPU i mod P)."""
    return list(range(rank, n, world))


def sharded_energies(system, configs: Sequence[Config], options=None, group=None) -> List[float]:
    """Evaluate `configs` across the ranks of a torch.distributed group (or locally) and return
    the full list on every rank.  Only scalars travel; no collective touches the matrices."""
    dist = _dist()
    on = dist is not None
    batch = getattr(system, "energyBatch", None)
    if batch is None:
        def batch(cs, opts=None):
            return [system.energy(c, opts or {}) for c in cs]
    if not on:
        return list(batch(list(configs), options))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_indices(len(configs), rank, world)
    local = batch([configs[i] for i in mine], options)
    gathered: List[Tuple[str, List[Tuple[int, float]]]] = [None] * world  # type: ignore
    dist.all_gather_object(gathered, (_configs_digest(configs), list(zip(mine, local))), group=group)
    digests = {d for d, _ in gathered}
    if len(digests) != 1:
        raise RuntimeError("sharded_energies: the ranks hold different candidate lists (unsynchronised RNG?); "
                           "every rank must pass the same configs -- see broadcast_from_rank0")
    out = [math.nan] * len(configs)
    for _, part in gathered:
        for i, e in part:
            out[i] = e
    return out


class ModelTuner:
    def __init__(self, model):
        self.system = model
        self.step = 0.3
        self.gridsize = 3
        self.logarithmicScale = False
        self.num_samples = 20
        self.meanFieldPrior: Dict[str, object] = {}
        self.group = None  # torch.distributed group used to shard candidates (None = default / local)

    # fluent setters (ModelTuner.scala:52-77)
    def setPrior(self, p):
        self.meanFieldPrior = dict(p)
        return self

    def setNumSamples(self, n: int):
        self.num_samples = int(n)
        return self

    def setLogScale(self, t: bool):
        self.logarithmicScale = bool(t)
        return self

    def setGridSize(self, s: int):
        self.gridsize = int(s)
        return self

    def setStepSize(self, s: float):
        self.step = float(s)
        return self

    def getGrid(self, initialConfig: Config) -> List[Config]:
        """ModelTuner.scala:79-103: per hyper-parameter `gridsize` values h0 - (i+1)*step, or
        h0 / exp((i+1)*step) on the log scale; cartesian product in key order."""
        keys = list(initialConfig.keys())
        if self.logarithmicScale:
            vecs = [[initialConfig[k] / math.exp((i + 1) * self.step) for i in range(self.gridsize)] for k in keys]
        else:
            vecs = [[initialConfig[k] - (i + 1) * self.step for i in range(self.gridsize)] for k in keys]
        return [dict(zip(keys, combo)) for combo in itertools.product(*vecs)]

    def getEnergyLandscape(self, initialConfig: Config, options=None, prior=None) -> List[Tuple[float, Config]]:
        """ModelTuner.scala:105-155.  A prior covering every key switches the grid to `num_samples`
        draws and adds -sum log prior(h); priors are objects with .sample() and .logPdf(x)."""
        prior = prior or {}
        use_prior = len(initialConfig) > 0 and all(k in prior for k in initialConfig)
        if use_prior:
            grid = [{k: float(prior[k].sample()) for k in prior} for _ in range(self.num_samples)]
            grid = broadcast_from_rank0(grid, self.group)
        else:
            grid = self.getGrid(initialConfig)
        energies = sharded_energies(self.system, grid, options, self.group)
        out = []
        for cfg, e in zip(grid, energies):
            pe = -sum(prior[k].logPdf(v) for k, v in cfg.items()) if use_prior else 0.0
            out.append((pe + e, cfg))
        return out


class GridSearch(ModelTuner):
    def optimize(self, initialConfig: Config, options=None):
        options = options or {}
        landscape = self.getEnergyLandscape(initialConfig, options, self.meanFieldPrior)
        # `.toMap` then `keys.min` (GridSearch.scala:47-49): ties collapse to the last config
        as_map = {e: cfg for e, cfg in landscape}
        optimum = min(as_map.keys())
        best = as_map[optimum]
        if options.get("persist") in ("true", "1"):
            self.system.persist(best)
        return self.system, best


class AbstractCSA(ModelTuner):
    MuSA, BA, M, MwVC, SA = "CSA-MuSA", "CSA-BA", "CSA-M", "CSA-MwVC", "SA"

    def __init__(self, model, rng: np.random.Generator | None = None):
        super().__init__(model)
        self.MAX_ITERATIONS = 10
        self.variant = AbstractCSA.MuSA
        self.iTemp = 1.0
        self.alpha = 0.05
        self.rng = rng if rng is not None else np.random.default_rng()

    def setVariant(self, v: str):
        self.variant = v
        return self

    def setMaxIterations(self, m: int):
        self.MAX_ITERATIONS = int(m)
        return self

    # ---- formulas (AbstractCSA.scala:268-327) -----------------------------------------------
    @staticmethod
    def couplingFactor(variant: str, landscape: Sequence[float], Tacc: float) -> float:
        # scala.math.exp / Double arithmetic: overflow is +Inf, Inf - Inf is NaN, nothing throws (:285-300)
        if variant in (AbstractCSA.MuSA, AbstractCSA.BA):
            return float(sum(np.float64(_exp(_div(-e, Tacc))) for e in landscape))
        if variant in (AbstractCSA.M, AbstractCSA.MwVC):
            return float(sum(np.float64(_exp(_div(e, Tacc))) for e in landscape))
        return 1.0

    @staticmethod
    def acceptanceProbability(variant: str, energy: float, oldEnergy: float, gamma: float, temperature: float) -> float:
        with np.errstate(all="ignore"):
            if variant == AbstractCSA.MuSA:
                x = np.float64(_exp(_div(-energy, temperature)))
                return _div(x, x + np.float64(gamma))
            if variant == AbstractCSA.BA:
                return float(np.float64(1.0) - np.float64(_div(_exp(_div(-oldEnergy, temperature)), gamma)))
            if variant in (AbstractCSA.M, AbstractCSA.MwVC):
                return _div(_exp(_div(oldEnergy, temperature)), gamma)
            return _div(gamma, np.float64(1.0) + np.float64(_exp(_div(np.float64(energy) - np.float64(oldEnergy),
                                                                      temperature))))

    @staticmethod
    def varianceDesired(variant: str, m: int) -> float:
        if variant in (AbstractCSA.MuSA, AbstractCSA.BA):
            return 0.99
        return 0.99 * (m - 1) / float(m) ** 2

    def acceptanceTemperature(self, initialTemp: float, k: int) -> float:
        return _div(initialTemp, math.log(k + 1.0))

    def mutationTemperature(self, initialTemp: float, k: int) -> float:
        return initialTemp / float(k)

    def mutate(self, config: Config, temperature: float) -> Config:
        """Cauchy(0, T) perturbation then abs (AbstractCSA.scala:112-130)."""
        return {k: abs(v + temperature * float(self.rng.standard_cauchy())) for k, v in config.items()}

    @staticmethod
    def _sample_variance(xs: Sequence[float]) -> float:
        """utils.getStats (CORE/utils/package.scala:110-143): unbiased variance."""
        n = len(xs)
        if n < 2:
            raise ValueError("To calculate stats size of data must be > 1")
        with np.errstate(all="ignore"):
            a = np.asarray(xs, dtype=np.float64)
            m = a.sum() / n
            return float(((a - m) * (a - m)).sum() / (n - 1))

    def performCSA(self, initialConfig: Config, options=None) -> List[Tuple[float, Config]]:
        options = options or {}
        # one process per GPU: every rank replays rank 0's random stream (mutations, accept draws)
        self.rng.bit_generator.state = broadcast_from_rank0(self.rng.bit_generator.state, self.group)
        sigmaD = self.varianceDesired(self.variant, int(self.gridsize ** len(initialConfig)))
        landscape = self.getEnergyLandscape(initialConfig, options, self.meanFieldPrior)
        gamma_init = self.couplingFactor(self.variant, [e for e, _ in landscape], self.iTemp)
        acceptanceProbs = [self.acceptanceProbability(self.variant, e, e, gamma_init, self.iTemp) for e, _ in landscape]
        use_prior = len(initialConfig) > 0 and all(k in self.meanFieldPrior for k in initialConfig)

        accTempPrev = self.iTemp
        for it in range(self.MAX_ITERATIONS, 0, -1):
            mutTemp = self.mutationTemperature(self.iTemp, it)
            if self.variant == AbstractCSA.MwVC:
                variance = self._sample_variance(acceptanceProbs)
                accTemp = accTempPrev * (1 - self.alpha) if variance < sigmaD else accTempPrev * (1 + self.alpha)
            else:
                accTemp = self.acceptanceTemperature(self.iTemp, it)
            maxEnergy = max(e for e, _ in landscape)
            with np.errstate(all="ignore"):
                shifted = [float(np.float64(e) - np.float64(maxEnergy)) for e, _ in landscape]
            coupling = self.couplingFactor(self.variant, shifted, accTemp)
            # candidates of one sweep are independent given the previous landscape: one batch
            proposals = [self.mutate(cfg, mutTemp) for _, cfg in landscape]
            energies = sharded_energies(self.system, proposals, options, self.group)
            new_landscape, probs = [], []
            for (old_e, old_cfg), new_cfg, e in zip(landscape, proposals, energies):
                pe = -sum(self.meanFieldPrior[k].logPdf(v) for k, v in new_cfg.items()) if use_prior else 0.0
                new_e = e + pe
                with np.errstate(all="ignore"):
                    d_new, d_old = float(np.float64(new_e) - np.float64(maxEnergy)), float(np.float64(old_e) - np.float64(maxEnergy))
                p = self.acceptanceProbability(self.variant, d_new, d_old, coupling, accTemp)
                if new_e < old_e or float(self.rng.random()) <= p:
                    new_landscape.append((new_e, new_cfg))
                else:
                    new_landscape.append((old_e, old_cfg))
                probs.append(p)
            landscape, acceptanceProbs, accTempPrev = new_landscape, probs, accTemp
        return landscape


class CoupledSimulatedAnnealing(AbstractCSA):
    def optimize(self, initialConfig: Config, options=None):
        options = options or {}
        as_map = {e: cfg for e, cfg in self.performCSA(initialConfig, options)}
        optimum = min(as_map.keys())
        best = as_map[optimum]
        if options.get("persist") in ("true", "1"):
            self.system.persist(best)
        return self.system, best


class ProbGPCommMachine(CoupledSimulatedAnnealing):
    """Build a GP committee from the energy landscape (CORE/optimization/ProbGPCommMachine.scala:36-229): every
    configuration of the landscape becomes a member with weight w_i; the committee is one GP with kernel
    sum_i w_i^2 k_i, noise sum_i w_i^2 noise_i and mean (sum_i w_i) m(x).  On the device the committee kernel is a
    DGP_KERNEL_COMPOSITE of the members (at most DGP_MAX_FACTORS base kernels)."""

    def __init__(self, model, rng: np.random.Generator | None = None):
        super().__init__(model, rng)
        self.policy = "CSA"
        self.baselinePolicy = "max"

    @property
    def _policy(self) -> str:
        return self.policy

    def setPolicy(self, p: str):
        self.policy = "CSA" if p in ("CSA", "Coupled Simulated Annealing") else "GS"
        return self

    def setBaseLinePolicy(self, p: str):
        if p in ("avg", "mean", "average"):
            self.baselinePolicy = "mean"
        elif p in ("min", "max"):
            self.baselinePolicy = p
        else:
            self.baselinePolicy = "mean"
        return self

    @staticmethod
    def calculateModelWeights(energyLandscape: Sequence[Tuple[float, Config]]) -> List[Tuple[float, Config]]:
        """:180-192: w_i = (1 - h_i / sum h) / (n - 1)."""
        h = [e for e, _ in energyLandscape]
        total = sum(h)
        alpha = 1.0 if len(h) == 1 else 1.0 / (len(h) - 1)
        return [((1.0 - e / total) * alpha, cfg) for e, cfg in energyLandscape]

    @staticmethod
    def calculateModelWeightsSigmoid(baselineMethod: str = "mean"):
        """:194-227: w_i = 1 / sum_{r != i} exp(-(h_r - h_i) / baseline) -- row i of the mask matrix is zero, so that
        term contributes exp(0) = 1 as well, exactly as `sum(exp((mask * h) *:* (-1/baseline)))` does."""
        def weights(energyLandscape: Sequence[Tuple[float, Config]]) -> List[Tuple[float, Config]]:
            h = [e for e, _ in energyLandscape]
            if baselineMethod == "min":
                baseline = min(h)
            elif baselineMethod == "max":
                baseline = max(h)
            else:
                baseline = sum(h) / len(h)
            out = []
            for i, (_, cfg) in enumerate(energyLandscape):
                tot = sum(math.exp((0.0 if r == i else h[r] - h[i]) * (-1.0 / baseline)) for r in range(len(h)))
                out.append((1.0 / tot, cfg))
            return out
        return weights

    def optimize(self, initialConfig: Config, options=None):
        import copy
        from .kernels import DecomposableCovariance, DiracKernel
        from .models import CudaGPRegression
        options = options or {}
        system = self.system
        blocked = list(system.covariance.blocked_hyper_parameters) + list(system.noiseModel.blocked_hyper_parameters)
        kernelParams, noiseParams = list(system.covariance.hyper_parameters), list(system.noiseModel.hyper_parameters)
        blockedState = {k: v for k, v in system._current_state.items() if k in blocked}
        if self.policy == "CSA":
            landscape = self.performCSA(initialConfig, options)
        else:
            landscape = self.getEnergyLandscape(initialConfig, options, self.meanFieldPrior)
        weights = [(w, {**cfg, **blockedState})
                   for w, cfg in self.calculateModelWeightsSigmoid(self.baselinePolicy)(landscape)]
        kernels, noise_level, wsum = [], 0.0, 0.0
        for w, conf in weights:
            k = copy.deepcopy(system.covariance)   # covariance.asPipe(conf): a fresh kernel fixed at this configuration
            k.setHyperParameters({h: conf[h] for h in k.effective_hyper_parameters if h in conf})
            n = copy.deepcopy(system.noiseModel)
            n.setHyperParameters({h: conf[h] for h in n.effective_hyper_parameters if h in conf})
            kernels.append(k)
            noise_level += w * w * abs(n.state["noiseLevel"])   # WeightedSumReducer(w^2) over Dirac terms
            wsum += w
        netKernel = DecomposableCovariance(kernels, [w * w for w, _ in weights])
        netNoise = DiracKernel(noise_level)
        base_mean = system.mean
        committee = CudaGPRegression(netKernel, netNoise, system.g, lambda x: wsum * base_mean(x), ctx=system.ctx)
        self.weights = weights
        if options.get("persist") in ("true", "1"):
            committee.persist(committee._current_state)
        return committee, committee._current_state


class GradBasedGlobalOptimizer(ModelTuner):
    """ML-II by plain gradient descent: h <- |h - step * g| (GradBasedGlobalOptimizer.scala:37-103).
    Note the reference descends along tr((aa' - K^-1) dK), which is -2 dE/dtheta."""

    def optimize(self, initialConfig: Config, options=None):
        options = options or {"tolerance": "0.0001", "step": "0.005", "maxIterations": "50"}
        tolerance, alpha, maxit = float(options["tolerance"]), float(options["step"]), int(options["maxIterations"])
        count, working = 1, dict(initialConfig)
        while True:
            gradient = self.system.gradEnergy(working)
            gradNorm = math.sqrt(sum(v * v for v in gradient.values()))
            # `working_solution.zip(gradient)` pairs the two maps positionally (:69)
            new = {}
            for (hyp, val), (_, gr) in zip(working.items(), gradient.items()):
                if gr == math.inf:
                    gr = 1.0
                elif gr == -math.inf:
                    gr = -1.0
                new[hyp] = abs(val - alpha * gr)  # NaN compares unequal to itself in the reference too
            working = new
            count += 1
            if not (count < maxit and gradNorm >= tolerance):
                break
        if options.get("persist") in ("true", "1"):
            self.system.persist(working)
        return self.system, working
