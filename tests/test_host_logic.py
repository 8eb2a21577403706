"""Host-side mirror of the reference API: hyper-parameter plumbing, grid construction, CSA formulas,
candidate sharding.  CPU only -- no compute call reaches the CUDA library here."""
import math
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

import dynaml_b200 as dg
from dynaml_b200 import _capi
from dynaml_b200 import optimization as opt
from oracle import gp_oracle as orc

ROOT = Path(__file__).resolve().parent.parent


def test_block_unblock_like_kernelspec():
    # KernelSpec.scala:12-32
    se = dg.SEKernel(band=1.0, h=1.0)
    hyp = se.hyper_parameters
    se.block(hyp[0])
    assert se.effective_hyper_parameters[0] == hyp[-1]
    se.block_all_hyper_parameters()
    assert se.effective_hyper_parameters == []
    se.block()
    assert len(se.effective_hyper_parameters) == 2


def test_set_hyper_parameters_requires_all_effective_keys():
    se = dg.SEKernel(1.0, 1.0)
    with pytest.raises(AssertionError):
        se.setHyperParameters({"bandwidth": 2.0})
    se.block("amplitude")
    se.setHyperParameters({"bandwidth": 2.0, "amplitude": 9.0})
    assert se.state == {"bandwidth": 2.0, "amplitude": 1.0}  # blocked keys are not overwritten


def test_matern_p_is_blocked_and_floored():
    # GenericMaternKernel.scala:42-53
    k = dg.GenericMaternKernel(1.5, 2)
    assert k.hyper_parameters == ["p", "l"] and k.effective_hyper_parameters == ["l"]
    k.block()
    k.setHyperParameters({"p": -2.7, "l": 0.5})
    assert k.state == {"p": 2.0, "l": 0.5}


def test_mahalanobis_names():
    k = dg.MahalanobisKernel([1.0, 2.0, 3.0], 1.5)
    assert k.hyper_parameters == ["MahalanobisAmplitude", "MahalanobisBandwidth_0", "MahalanobisBandwidth_1",
                                  "MahalanobisBandwidth_2"]
    assert k._device_hyp(k.state) == [1.5, 1.0, 2.0, 3.0]


def test_additive_covariance_naming_and_routing():
    # KernelSpec.scala:87-129; AdditiveCovariance.scala:32-58
    se, nz = dg.SEKernel(1.0, 2.0), dg.DiracKernel(0.5)
    nz.block_all_hyper_parameters()
    add = se + nz
    fid, sid = str(se).split(".")[-1], str(nz).split(".")[-1]
    assert add.hyper_parameters == [f"{fid}/bandwidth", f"{fid}/amplitude", f"{sid}/noiseLevel"]
    assert add.effective_hyper_parameters == [f"{fid}/bandwidth", f"{fid}/amplitude"]
    add.setHyperParameters({f"{fid}/bandwidth": 3.0, f"{fid}/amplitude": 4.0})
    assert se.state == {"bandwidth": 3.0, "amplitude": 4.0} and nz.state == {"noiseLevel": 0.5}


def test_kernel_algebra_reduces_to_sum_of_products():
    # LocalScalarKernel.scala:53-91: +, *, * c; the device evaluates sum_t coef_t prod_f leaf_f^e_tf
    se, per, rbf, mat = dg.SEKernel(1.5, 1.2), dg.PeriodicKernel(1.0, 1.3, 0.2), dg.RBFKernel(2.0), \
        dg.GenericMaternKernel(1.1, 2)
    k = (se + per * 2.0) * (rbf + mat) * 0.5
    assert [type(x).__name__ for x in k._leaves()] == ["SEKernel", "PeriodicKernel", "RBFKernel", "GenericMaternKernel"]
    terms = sorted((c, tuple(sorted(m.items()))) for c, m in k._terms())
    assert terms == sorted([(0.5, ((0, 1), (2, 1))), (0.5, ((0, 1), (3, 1))), (1.0, ((1, 1), (2, 1))),
                            (1.0, ((1, 1), (3, 1)))])
    # names nest one "<Class@hash>/" per combinator, in leaf order; Matern's p stays blocked through the nesting
    assert [h.split("/")[-1] for h in k.hyper_parameters] == ["bandwidth", "amplitude", "sigma", "lengthscale",
                                                              "frequency", "bandwidth", "p", "l"]
    assert [h.split("/")[-1] for h in k.effective_hyper_parameters][-1] == "l" and \
        "p" not in [h.split("/")[-1] for h in k.effective_hyper_parameters]
    k.setHyperParameters({h: 3.0 for h in k.effective_hyper_parameters})
    assert se.state == {"bandwidth": 3.0, "amplitude": 3.0} and rbf.state == {"bandwidth": 3.0}
    assert mat.state["l"] == 3.0 and mat.state["p"] == 2.0
    spec, prog = k._spec(noise=0.1)
    assert spec.kind == _capi.KERNEL_COMPOSITE and prog[0] == 4 and prog[1] == 4
    assert _capi.load_library().dgp_spec_grad_len(spec) == 2 + 3 + 1 + 2 + 1
    s1, _ = dg.SEKernel(1.0, 1.0)._spec()
    assert _capi.load_library().dgp_spec_grad_len(s1) == 3
    with pytest.raises(ValueError):
        _ = se * -1.0  # "Multiplicative constant applied on a kernel must be positive!"


def test_cov_plus_dirac_keeps_the_single_kernel_noise_path():
    se, nz = dg.SEKernel(1.0, 2.0), dg.DiracKernel(0.5)
    spec, _ = (se + nz)._spec()
    assert spec.kind == _capi.KERNEL_SE and spec.noise == 0.5
    spec2, _ = ((se * 2.0) + nz)._spec()
    assert spec2.kind == _capi.KERNEL_COMPOSITE and spec2.noise == 0.0


def test_unsupported_kernel_has_no_cpu_fallback():
    class Cauchy(dg.LocalScalarKernel):
        hyper_parameters = ["sigma"]

    with pytest.raises(NotImplementedError):
        Cauchy()._spec()


def test_dirac_builder_override_is_signed():
    # DiracKernel.scala:55-58
    K = dg.DiracKernel(-0.3).buildKernelMatrix([[0.0], [1.0]], 2).getKernelMatrix()
    assert np.array_equal(K, -0.3 * np.eye(2))


class FakeSystem:
    """A GloballyOptimizable whose energy is a known function (no GPU)."""

    def __init__(self):
        self.calls = []
        self.persisted = None

    def energy(self, h, options=None):
        self.calls.append(dict(h))
        return (h["a"] - 1.0) ** 2 + (h["b"] - 0.5) ** 2

    def energyBatch(self, hs, options=None):
        return [self.energy(h, options) for h in hs]

    def persist(self, h):
        self.persisted = dict(h)


@pytest.mark.parametrize("log_scale", [False, True])
def test_get_grid_matches_oracle(log_scale):
    t = dg.GridSearch(FakeSystem()).setGridSize(4).setStepSize(0.2).setLogScale(log_scale)
    init = {"a": 2.0, "b": 1.5}
    grid = t.getGrid(init)
    assert grid == orc.get_grid(init, 4, 0.2, log_scale)
    assert len(grid) == 16
    # ModelTuner.scala:83-86
    first = 2.0 / math.exp(0.2) if log_scale else 2.0 - 0.2
    assert grid[0]["a"] == first


def test_grid_search_picks_minimum_and_persists():
    sys_ = FakeSystem()
    t = dg.GridSearch(sys_).setGridSize(5).setStepSize(0.25)
    _, best = t.optimize({"a": 2.0, "b": 1.5}, {"persist": "true"})
    e, cfg = orc.grid_search(lambda c: FakeSystem().energy(c), {"a": 2.0, "b": 1.5}, 5, 0.25, False)
    assert best == cfg and sys_.persisted == cfg
    assert len(sys_.calls) == 25


@pytest.mark.parametrize("variant", ["CSA-MuSA", "CSA-BA", "CSA-M", "CSA-MwVC", "SA"])
def test_csa_formulas_match_oracle(variant):
    land = [3.0, 1.0, 2.5, 0.2]
    t = 0.7
    g = dg.CoupledSimulatedAnnealing.couplingFactor(variant, land, t)
    assert g == pytest.approx(orc.csa_coupling(variant, land, t), rel=1e-15)
    p = dg.CoupledSimulatedAnnealing.acceptanceProbability(variant, 1.2, 2.0, g, t)
    assert p == pytest.approx(orc.csa_acceptance(variant, 1.2, 2.0, g, t), rel=1e-15)


def test_csa_runs_and_never_worsens_best():
    sys_ = FakeSystem()
    csa = dg.CoupledSimulatedAnnealing(sys_, rng=np.random.default_rng(0)).setGridSize(3).setStepSize(0.3)
    csa.setMaxIterations(6).setVariant(dg.CoupledSimulatedAnnealing.MwVC)
    init = {"a": 2.0, "b": 1.5}
    start = min(e for e, _ in csa.getEnergyLandscape(init))
    _, best = csa.optimize(init)
    assert sys_.energy(best) <= start + 1e-12
    assert all(v >= 0 for v in best.values())  # mutate takes abs()
    # temperatures (AbstractCSA.scala:132-136)
    assert csa.acceptanceTemperature(1.0, 3) == 1.0 / math.log(4.0) and csa.mutationTemperature(1.0, 4) == 0.25


def test_grad_descent_step_rule():
    class G(FakeSystem):
        def gradEnergy(self, h):
            return {"a": 2 * (h["a"] - 1.0), "b": 2 * (h["b"] - 0.5)}

    _, sol = dg.GradBasedGlobalOptimizer(G()).optimize({"a": 2.0, "b": 1.5},
                                                       {"tolerance": "1e-9", "step": "0.1", "maxIterations": "200"})
    assert sol["a"] == pytest.approx(1.0, abs=1e-6) and sol["b"] == pytest.approx(0.5, abs=1e-6)


def test_shard_indices_cover_everything_once():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in opt.shard_indices(n, r, world))
            assert got == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys_ = FakeSystem()
        t = dg.GridSearch(sys_).setGridSize(4).setStepSize(0.2)
        land = t.getEnergyLandscape({"a": 2.0, "b": 1.5})
        _, best = t.optimize({"a": 2.0, "b": 1.5})
        # every rank sees the full landscape but evaluated only its own shard (twice: landscape + optimize)
        q.put((rank, [e for e, _ in land], best, len(sys_.calls)))
    finally:
        dist.destroy_process_group()


def test_candidates_shard_across_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    serial = [e for e, _ in dg.GridSearch(FakeSystem()).setGridSize(4).setStepSize(0.2)
              .getEnergyLandscape({"a": 2.0, "b": 1.5})]
    assert res[0][1] == serial and res[1][1] == serial
    assert res[0][2] == res[1][2]
    assert res[0][3] == 16 and res[1][3] == 16  # 8 of the 16 candidates, two sweeps


class _HugeEnergySystem(FakeSystem):
    """Energies of the size a GP NLL has at N = 32768 (|E| ~ 3e4), one candidate non-PD (+Inf)."""

    def __init__(self, sign):
        super().__init__()
        self.sign = sign

    def energy(self, h, options=None):
        self.calls.append(dict(h))
        if len(self.calls) % 7 == 3:
            return math.inf
        return self.sign * 1e4 * (1.0 + (h["a"] - 1.0) ** 2 + (h["b"] - 0.5) ** 2)


@pytest.mark.parametrize("variant", ["CSA-MuSA", "CSA-BA", "CSA-M", "CSA-MwVC", "SA"])
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_csa_runs_at_realistic_energy_magnitudes(variant, sign):
    """AbstractCSA.scala:285-325 is plain Double arithmetic: exp overflows to Inf, Inf/Inf is NaN, nothing throws
    (math.exp would raise OverflowError above ~709).  Every variant must get through a whole annealing run."""
    sys_ = _HugeEnergySystem(sign)
    csa = dg.CoupledSimulatedAnnealing(sys_, rng=np.random.default_rng(3)).setGridSize(3).setStepSize(0.2)
    csa.setVariant(variant).setMaxIterations(4)
    land = csa.performCSA({"a": 2.0, "b": 1.5})
    assert len(land) == 9
    _, best = csa.optimize({"a": 2.0, "b": 1.5})
    assert set(best) == {"a", "b"}
    A = dg.CoupledSimulatedAnnealing
    for e in (3e4, -3e4, 5.0, math.inf):
        g = A.couplingFactor(variant, [e, e + 800.0, e - 900.0], 1.0)
        A.acceptanceProbability(variant, e, e - 1.0, g, 1.0)      # may be NaN / Inf, must not raise
    assert A.couplingFactor("CSA-MuSA", [-3e4], 1.0) == math.inf
    assert A.couplingFactor("CSA-MuSA", [3e4], 1.0) == 0.0
    assert math.isnan(A.acceptanceProbability("CSA-MuSA", 3e4, 3e4, 0.0, 1.0))   # 0 / (0 + 0)


def _gloo_csa_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank starts from its OWN random stream, as the reference's unseeded RNGs would
        csa = dg.CoupledSimulatedAnnealing(FakeSystem(), rng=np.random.default_rng(1000 + rank))
        csa.setGridSize(3).setStepSize(0.2).setMaxIterations(3)
        land = csa.performCSA({"a": 2.0, "b": 1.5})
        ok = all(abs(e - ((c["a"] - 1.0) ** 2 + (c["b"] - 0.5) ** 2)) < 1e-15 for e, c in land)
        mismatch = None
        try:   # ranks passing different candidate lists must be caught, not silently mis-paired
            from dynaml_b200.optimization import sharded_energies
            sharded_energies(FakeSystem(), [{"a": 1.0 + rank, "b": 0.0}, {"a": 0.0, "b": 1.0}])
        except RuntimeError as ex:
            mismatch = str(ex)
        q.put((rank, land, ok, mismatch))
    finally:
        dist.destroy_process_group()


def test_csa_is_consistent_across_two_gloo_ranks():
    """One process per GPU: proposals and accept draws are rank 0's on every rank, so each (energy, config) pair of
    the landscape belongs together and all ranks end in the same state."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_csa_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                 # identical landscapes
    assert res[0][2] and res[1][2]                # every energy is the energy of ITS config
    assert res[0][3] and "different candidate lists" in res[0][3]


# ---- blocked containers (CORE/algebra/PartitionedMatrix.scala, PartitionedVector.scala) ---------------------------------
def test_partitioned_containers():
    from dynaml_b200.algebra import LowerTriPartitionedMatrix, PartitionedMatrix, PartitionedPSDMatrix, PartitionedVector
    rng = np.random.default_rng(0)
    A = rng.standard_normal((7, 7))
    S = A @ A.T
    cuts = [(0, 3), (3, 6), (6, 7)]
    lower = [((i, j), S[a:b, c:d]) for i, (a, b) in enumerate(cuts) for j, (c, d) in enumerate(cuts) if i >= j]
    P = PartitionedPSDMatrix(lower, 7, 7, 3, 3)
    assert (P.rows, P.cols, P.rowBlocks, P.colBlocks) == (7, 7, 3, 3)
    assert len(P._underlyingdata) == 6 and len(P._data) == 9          # upper blocks are the transposes (:248-259)
    assert np.array_equal(P.toBreezeMatrix(), S)
    sub = P(range(1, 3), range(0, 2))                                 # re-indexed from 0 (:50-55)
    assert (sub.rowBlocks, sub.colBlocks) == (2, 2) and np.array_equal(sub.toBreezeMatrix(), S[3:7, 0:6])
    assert np.array_equal(P.t.toBreezeMatrix(), S.T)
    L = np.linalg.cholesky(S)
    T = LowerTriPartitionedMatrix([((i, j), L[a:b, c:d]) for i, (a, b) in enumerate(cuts) for j, (c, d) in enumerate(cuts)
                                   if i >= j], 7, 7, 3, 3)
    assert np.array_equal(T.toBreezeMatrix(), L) and T._blocks()[(0, 2)].shape == (3, 1)
    M = PartitionedMatrix([((0, 0), np.ones((2, 3))), ((1, 0), 2 * np.ones((1, 3)))])
    assert (M.rows, M.cols, M.rowBlocks, M.colBlocks) == (3, 3, 2, 1)
    v = PartitionedVector.fromDense(np.arange(7.0), 3)
    assert v.rowBlocks == 3 and v.rows == 7 and [x.size for _, x in v._data] == [3, 3, 1]
    assert np.array_equal(v.toBreezeVector(), np.arange(7.0)) and np.array_equal(v(range(1, 3)).toBreezeVector(), np.arange(3.0, 7.0))


# ---- 2D block-cyclic layout of the distributed Cholesky (host-only entry point of the C ABI) ----------
@pytest.mark.parametrize("n,tile,pr,pc", [(131072, 512, 2, 4), (131072, 512, 1, 1), (10000, 256, 2, 2),
                                          (3000, 256, 1, 2), (700, 256, 2, 4)])
def test_block_cyclic_plan_partitions_the_lower_triangle(n, tile, pr, pc):
    from dynaml_b200 import _capi
    plans = [_capi.dist_plan(n, tile, pr, pc, r) for r in range(pr * pc)]
    nt = -(-n // tile)
    assert all(p["tiles"] == nt for p in plans)
    # every lower tile has exactly one owner: (I mod Pr, J mod Pc)
    owned = [0] * (pr * pc)
    for i in range(nt):
        for j in range(i + 1):
            owned[(i % pr) * pc + (j % pc)] += 1
    assert [p["owned_lower_tiles"] for p in plans] == owned
    assert sum(owned) == nt * (nt + 1) // 2
    for r, p in enumerate(plans):
        assert p["local_rows"] == len(range(r // pc, nt, pr)) and p["local_cols"] == len(range(r % pc, nt, pc))
        # packed lower storage (dist.cu, Store): chunks of 8 local tile columns, each from the first local row its first
        # column needs; about half of the local rectangle on a large grid, never more than all of it
        mt, ntc, prow, pcol = p["local_rows"], p["local_cols"], r // pc, r % pc
        want = 0
        for c0 in range(0, ntc, 8):
            J = c0 * pc + pcol
            first = 0 if J <= prow else min(mt, -(-(J - prow) // pr))
            want += (mt - first) * min(8, ntc - c0)
        assert p["local_bytes"] == want * tile * tile * 8 <= mt * ntc * tile * tile * 8
        if nt >= 64:
            assert p["local_bytes"] <= 0.57 * mt * ntc * tile * tile * 8


def test_gaussian_scaling_and_regression_metrics():
    from dynaml_b200 import workflow as wf
    rng = np.random.default_rng(0)
    X = rng.normal(3.0, 2.0, size=(40, 3))
    y = rng.normal(-1.0, 5.0, size=40)
    tr, te, (mx, sx, my, sy) = wf.gaussianScalingTrainTest(list(zip(X[:30], y[:30])), list(zip(X[30:], y[30:])))
    Xs = np.array([t[0] for t in tr])
    assert np.allclose(Xs.mean(0), 0.0, atol=1e-12) and np.allclose(Xs.std(0, ddof=1), 1.0)
    assert np.allclose(np.array([t[0] for t in te]), (X[30:] - mx) / sx)
    m = wf.RegressionMetrics([(1.0, 1.5), (2.0, 1.0), (3.0, 3.5)])
    assert m.mae == pytest.approx((0.5 + 1.0 + 0.5) / 3) and m.rmse == pytest.approx(math.sqrt((0.25 + 1 + 0.25) / 3))
    assert m.Rsq == pytest.approx(1 - 1.5 / ((1.5 - 2) ** 2 + (1 - 2) ** 2 + (3.5 - 2) ** 2))


# ---- consumers of `energy` ----------------------------------------------------------------
class _QuadraticSystem:
    """GloballyOptimizable stand-in: energy = 0.5 (h - mu)' P (h - mu), i.e. a Gaussian posterior N(mu, P^-1)."""

    def __init__(self):
        self.mu = np.array([1.5, 0.7])
        self.P = np.array([[4.0, 1.0], [1.0, 9.0]])
        self._current_state = {"a": 1.0, "b": 1.0}
        self.calls = 0

    def energy(self, h, options=None):
        self.calls += 1
        v = np.array([h["a"], h["b"]]) - self.mu
        return 0.5 * float(v @ self.P @ v)


class _FlatPrior:
    def logPdf(self, x):
        return 0.0


def test_adaptive_mcmc_samples_the_posterior_and_follows_the_reference_formulas():
    from dynaml_b200.mcmc import AdaptiveHyperParameterMCMC, HyperParameterMCMC, HyperParameterSCAM
    sys_ = _QuadraticSystem()
    prior = {"a": _FlatPrior(), "b": _FlatPrior()}
    mc = AdaptiveHyperParameterMCMC(sys_, prior, burnIn=300, rng=np.random.default_rng(0))
    draws = np.array([[s["a"], s["b"]] for s in mc.iid(3000)])
    assert np.abs(draws.mean(axis=0) - sys_.mu).max() < 0.08
    assert np.abs(np.cov(draws.T) - np.linalg.inv(sys_.P)).max() < 0.05
    assert 0.0 < mc.sampleAcceptenceRate <= 1.0 and mc._count > 3300
    # HyperParameterMCMC.scala:221-227 with the chain's own count / sigma
    adj = mc.count / (mc.count - 1)
    want = mc.sigma * ((1 - 0.05) * 2.38) ** 2 * adj / 2.0 + np.eye(2) * (0.1 * 0.05) ** 2 / 2.0
    assert np.allclose(mc.getExplorationVar(), want, rtol=0, atol=0)
    # :278-287 against a plain restatement
    m, s, x = mc.mean.copy(), mc.sigma.copy(), np.array([0.3, -0.2])
    mc.updateMoments(x)
    mnew = m + (x - m) / (mc.count + 1)
    assert np.allclose(mc.mean, mnew) and np.allclose(
        mc.sigma, s + np.outer(m, m) - np.outer(mnew, mnew) + (np.outer(x, x) - s - np.outer(m, m)) / (mc.count + 1))
    # the first 2 d proposals use eye * 0.01 / d
    fresh = AdaptiveHyperParameterMCMC(_QuadraticSystem(), prior, burnIn=0, rng=np.random.default_rng(1))
    assert np.array_equal(fresh.getExplorationVar(), np.eye(2) * 0.01 / 2.0)
    assert fresh._previous_sample == {"a": 1.0, "b": 1.0} and np.array_equal(fresh.sigma, np.ones((2, 2)))
    # SCAM: eye * 25 for count <= 10, then the scaled diagonal (:337-343)
    scam = HyperParameterSCAM(_QuadraticSystem(), prior, burnIn=0, rng=np.random.default_rng(2))
    assert np.array_equal(scam.getExplorationVar(), np.eye(2) * 25.0)
    scam.iid(30)
    adj = scam.count / (scam.count - 1)
    assert np.allclose(scam.getExplorationVar(), np.diag((np.diag(scam.sigma) * adj + 0.05) * 2.4 * 2.4))

    class _Proposal:          # vanilla MCMC draws candidates from a fixed random variable (:71-74)
        def __init__(self, rng):
            self.rng = rng

        def draw(self):
            return np.array([1.5, 0.7]) + 0.6 * self.rng.standard_normal(2)
    van = HyperParameterMCMC(_QuadraticSystem(), prior, _Proposal(np.random.default_rng(3)), burnIn=50,
                             rng=np.random.default_rng(4))
    sig0 = van.sigma.copy()
    van.iid(20)
    assert np.array_equal(van.sigma, sig0)  # never adapted


def test_committee_weights_match_the_mask_matrix_formula():
    # ProbGPCommMachine.scala:194-227 restated with the explicit mask matrices
    h = np.array([10.0, 12.5, 9.0, 30.0])
    landscape = [(float(e), {"i": float(i)}) for i, e in enumerate(h)]
    for method, baseline in (("mean", h.mean()), ("min", h.min()), ("max", h.max()), ("other", h.mean())):
        want = []
        for i in range(len(h)):
            mask = np.zeros((len(h), len(h)))
            for r in range(len(h)):
                for s in range(len(h)):
                    mask[r, s] = 0.0 if r == i else (-1.0 if s == i else (1.0 if r == s else 0.0))
            want.append(1.0 / np.exp((mask @ h) * (-1.0 / baseline)).sum())
        got = opt.ProbGPCommMachine.calculateModelWeightsSigmoid(method)(landscape)
        assert np.allclose([w for w, _ in got], want, rtol=1e-15)
        assert [c for _, c in got] == [c for _, c in landscape]
    lin = opt.ProbGPCommMachine.calculateModelWeights(landscape)   # :180-192
    assert np.allclose([w for w, _ in lin], (1.0 - h / h.sum()) / 3.0)
    m = opt.ProbGPCommMachine.__new__(opt.ProbGPCommMachine)
    m.policy, m.baselinePolicy = "CSA", "max"
    assert m.setPolicy("GS")._policy == "GS" and m.setPolicy("Coupled Simulated Annealing")._policy == "CSA"
    assert m.setBaseLinePolicy("average").baselinePolicy == "mean" and m.setBaseLinePolicy("zzz").baselinePolicy == "mean"


def test_decomposable_covariance_naming_and_terms():
    a, b = dg.SEKernel(1.0, 2.0), dg.RBFKernel(3.0)
    k = dg.DecomposableCovariance([a, b], [0.25, 0.75])
    ia, ib = str(a).split(".")[-1], str(b).split(".")[-1]
    assert k.hyper_parameters == [f"{ia}/bandwidth", f"{ia}/amplitude", f"{ib}/bandwidth"]
    assert k._terms() == [(0.25, {0: 1}), (0.75, {1: 1})]
    k.setHyperParameters({f"{ia}/bandwidth": 5.0, f"{ia}/amplitude": 6.0, f"{ib}/bandwidth": 7.0})
    assert a.state == {"bandwidth": 5.0, "amplitude": 6.0} and b.state == {"bandwidth": 7.0}
    assert k.state[f"{ib}/bandwidth"] == 7.0
