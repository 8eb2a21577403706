"""The oracle against every known-answer value the reference's own tests hold for this path
(SURVEY.md 8c), plus self-consistency of the restatement.  CPU only."""
import json
import math
from pathlib import Path

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import gp_oracle as orc

GOLD = Path(__file__).parent / "golden"
KATS = json.loads((GOLD / "kats.json").read_text())


def test_gpspec_energy_identity():
    # GPSpec.scala:13-38: Dirac(0.5)+Dirac(0.5) on 2 points, mean v(0): K = I, residual 0 => log 2pi
    X = np.array([[0.0], [1.0]])
    y = np.array([0.0, 1.0])
    e = orc.energy(orc.Kernel("dirac", {"noiseLevel": 0.5}), 0.5, X, y, mean=X[:, 0])
    assert abs(e - KATS["gp_energy_identity"]["value"]) < KATS["gp_energy_identity"]["eps"]
    assert e == pytest.approx(math.log(2 * math.pi), abs=1e-15)


def test_gpspec_posterior_exact():
    # GPSpec.scala:40-76: one training point, SE(1,1), noise 0: exact == in the reference
    k = orc.Kernel("se", {"bandwidth": 1.0, "amplitude": 1.0})
    mu, cov = orc.predictive(k, 0.0, np.array([[1.0]]), np.array([1.0]), np.array([[2.0]]))
    assert mu[0] == KATS["gp_posterior_mean"]["value"]
    assert cov[0, 0] == KATS["gp_posterior_var"]["value"]
    lo, hi = orc.error_bars(mu, cov, 1.0)
    assert lo[0] == KATS["gp_bar_lo"]["value"] and hi[0] == KATS["gp_bar_hi"]["value"]


def test_kernelspec_se_value():
    # KernelSpec.scala:34-58
    k = orc.Kernel("se", {"bandwidth": 1.0, "amplitude": 1.0})
    v = k.value(np.array([[1.0]]), np.array([[0.0]]))[0, 0]
    assert abs(v - KATS["se_value_1_0"]["value"]) < KATS["se_value_1_0"]["eps"]


def test_kernelspec_laplace_and_cauchy_values():
    # KernelSpec.scala:34-58: laplaceKernel(be=1)(1, 0) and cauchyKernel(1.0)(1, 1.5)
    lap = orc.Kernel("laplacian", {"beta": 1.0}).value(np.array([[1.0]]), np.array([[0.0]]))[0, 0]
    cau = orc.Kernel("cauchy", {"sigma": 1.0}).value(np.array([[1.0]]), np.array([[1.5]]))[0, 0]
    assert abs(lap - KATS["laplace_value_1_0"]["value"]) < KATS["laplace_value_1_0"]["eps"]
    assert abs(cau - KATS["cauchy_value_1_1p5"]["value"]) < KATS["cauchy_value_1_1p5"]["eps"]


def test_kernelspec_builders_layout():
    # KernelSpec.scala:181-231: eval2(x,y) = 1/(1+(x-y)^2) on the points 0,1 -> [[1,.5],[.5,1]], cross with {0} -> [1,.5]
    class Eval2(orc.Kernel):
        def __init__(self):
            super().__init__("rbf", {"bandwidth": 1.0})

        def value(self, X, Y):
            return 1.0 / (1.0 + orc._sqdist(X, Y))

    X = np.array([[0.0], [1.0]])
    K = orc.build_kernel_matrix(Eval2(), X)
    assert K.shape == (2, 2) and np.array_equal(K, np.array([[1.0, 0.5], [0.5, 1.0]]))
    C = orc.cross_kernel_matrix(Eval2(), X, np.array([[0.0]]))
    assert C.shape == (2, 1) and np.array_equal(C[:, 0], np.array([1.0, 0.5]))
    # KernelSpec.scala:205-229: the partitioned builders with one element per block
    blocks, rows, cols, rb, cb = orc.build_partitioned_kernel_matrix(Eval2(), X, 1, 1)
    assert (rows, cols, rb, cb) == (2, 2, 2, 2)                                   # k4.rows/cols/rowBlocks/colBlocks
    assert sorted(blocks) == [(0, 0), (1, 0), (1, 1)]                             # block row >= block column only
    assert all(np.array_equal(b, [[1.0]] if i == j else [[0.5]]) for (i, j), b in blocks.items())
    blocks, rows, cols, rb, cb = orc.cross_partitioned_kernel_matrix(Eval2(), X, np.array([[0.0]]), 1, 1)
    assert (rows, cols, rb, cb) == (2, 1, 2, 1)                                   # k5
    assert all(np.array_equal(b, [[1.0]] if i == j else [[0.5]]) for (i, j), b in blocks.items())


def test_partitioned_builder_blocks_are_slices_of_the_dense_matrix():
    """Ragged last block, noise on the diagonal blocks only (distinct points), Dirac on duplicates across blocks."""
    X, _, Xs = orc.synthetic(23, 3, seed=1, m=7)
    cov, nz = orc.Kernel("matern", {"p": 2.0, "l": 1.7}), orc.dirac(0.3)
    K = orc.build_kernel_matrix(cov, X, nz)
    blocks, rows, cols, rb, cb = orc.build_partitioned_kernel_matrix(cov, X, 5, 5, nz)
    assert (rows, cols, rb, cb) == (23, 23, 5, 5) and len(blocks) == 15
    for (i, j), b in blocks.items():
        assert np.array_equal(b, K[5 * i:5 * i + 5, 5 * j:5 * j + 5])
    X[17] = X[2]                                   # a duplicate in another block: the Dirac term fires off the diagonal
    blocks, *_ = orc.build_partitioned_kernel_matrix(cov, X, 5, 5, nz)
    assert blocks[(3, 0)][2, 2] == orc.build_kernel_matrix(cov, X, nz)[17, 2] == 1.0 + 0.3
    C = orc.cross_kernel_matrix(cov, X, Xs)
    blocks, rows, cols, rb, cb = orc.cross_partitioned_kernel_matrix(cov, X, Xs, 10, 4)
    assert (rows, cols, rb, cb) == (23, 7, 3, 2)
    for (i, j), b in blocks.items():
        assert np.array_equal(b, C[10 * i:10 * i + 10, 4 * j:4 * j + 4])


def test_partitioned_cholesky_spec():
    # PartitionedLinearAlgebraSpec.scala:14-46: 1000x1000 diagonal (i -> i+1?) in 5 blocks of 200
    n = 1000
    A = np.diag(np.arange(1, n + 1, dtype=float) ** 2)
    L = orc.bcholesky(A, 200)
    assert np.abs(L - np.diag(np.arange(1, n + 1, dtype=float))).max() <= KATS["bcholesky_diag_tol"]["value"]


def test_blocked_equals_dense():
    rng = np.random.default_rng(3)
    n = 350
    G = rng.standard_normal((n, n))
    A = G @ G.T + n * np.eye(n)
    Ld = sla.cholesky(A, lower=True)
    for blk in (100, 175, 350, 1000):
        Lb = orc.bcholesky(A, blk)
        assert np.abs(Lb - Ld).max() / np.abs(Ld).max() < 1e-13
        y = rng.standard_normal(n)
        z = orc.lower_solve_blocked(Lb, y, blk)
        assert np.allclose(z, sla.solve_triangular(Ld, y, lower=True), rtol=1e-11, atol=1e-13)
        x = orc.upper_solve_blocked(Lb.T, z, blk)
        assert np.allclose(x, np.linalg.solve(A, y), rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("kind,hyp", [
    ("se", {"bandwidth": 1.3, "amplitude": 0.9}),
    ("rbf", {"bandwidth": 1.3}),
    ("ard", {"MahalanobisAmplitude": 1.2, "MahalanobisBandwidth_0": 1.0, "MahalanobisBandwidth_1": 2.0,
             "MahalanobisBandwidth_2": 1.5}),
    ("matern", {"p": 1.0, "l": 1.4}),
    ("matern", {"p": 2.0, "l": 1.4}),
    ("matern", {"p": 3.0, "l": 1.4}),
    ("cauchy", {"sigma": 1.7}),
    ("laplacian", {"beta": 2.2}),
    ("ratquad", {"mu": 1.5, "l": 1.3}),
    ("tstudent", {"d": 1.6}),
])
def test_gradient_formulas_match_finite_differences(kind, hyp):
    """gradEnergy returns tr((aa' - K^-1) dK) = -2 d/dtheta of the *textbook* NLL."""
    X, y, _ = orc.synthetic(60, 3, seed=11)
    noise = 0.4

    def textbook(h, nz):
        K = orc.build_kernel_matrix(orc.Kernel(kind, h, "exact"), X, orc.dirac(nz))
        L = sla.cholesky(K, lower=True)
        a = sla.cho_solve((L, True), y)
        return 0.5 * y @ a + np.log(np.diag(L)).sum()

    g = orc.grad_energy(orc.Kernel(kind, hyp, "exact"), noise, X, y)
    for name in g:
        eps = 1e-6
        if name == "noiseLevel":
            fd = (textbook(hyp, noise + eps) - textbook(hyp, noise - eps)) / (2 * eps)
        else:
            hp, hm = dict(hyp), dict(hyp)
            hp[name] += eps
            hm[name] -= eps
            fd = (textbook(hp, noise) - textbook(hm, noise)) / (2 * eps)
        assert g[name] == pytest.approx(-2.0 * fd, rel=2e-5, abs=1e-6), name


def test_composite_gradients_match_finite_differences():
    """Sum / Prod / Scaled (AdditiveCovariance.scala:77-84, MultiplicativeCovariance.scala:68-77,
    LocalScalarKernel.scala:81-85) against central differences of the textbook NLL."""
    X, y, _ = orc.synthetic(50, 3, seed=13)
    noise = 0.4

    def build(h):
        se = orc.Kernel("se", {"bandwidth": h[0], "amplitude": h[1]})
        ca = orc.Kernel("cauchy", {"sigma": h[2]})
        ma = orc.Kernel("matern", {"p": 2.0, "l": h[3]})
        return orc.Sum(orc.Prod(se, ca), orc.Scaled(ma, 0.6))

    def textbook(h):
        K = orc.build_kernel_matrix(build(h), X, orc.dirac(noise))
        L = sla.cholesky(K, lower=True)
        a = sla.cho_solve((L, True), y)
        return 0.5 * y @ a + np.log(np.diag(L)).sum()

    h0 = [1.4, 1.1, 2.0, 1.7]
    cov = build(h0)
    assert cov.names() == ["k1/k1/bandwidth", "k1/k1/amplitude", "k1/k2/sigma", "k2/p", "k2/l"]
    assert cov.effective() == ["k1/k1/bandwidth", "k1/k1/amplitude", "k1/k2/sigma", "k2/l"]
    g = orc.grad_energy(cov, noise, X, y)
    for i, name in enumerate(cov.effective()):
        hp, hm = list(h0), list(h0)
        hp[i] += 1e-6
        hm[i] -= 1e-6
        fd = (textbook(hp) - textbook(hm)) / 2e-6
        assert g[name] == pytest.approx(-2.0 * fd, rel=2e-5, abs=1e-6), name


def test_reference_energy_has_half_logdet():
    # the log-det term is sum log L_ii, not 2 * that
    X, y, _ = orc.synthetic(40, 2, seed=5)
    k = orc.Kernel("se", {"bandwidth": 1.0, "amplitude": 1.0})
    K = orc.build_kernel_matrix(k, X, orc.dirac(0.3))
    sign, logdet = np.linalg.slogdet(K)
    a = np.linalg.solve(K, y)
    expect = 0.5 * (y @ a + 0.5 * logdet + len(y) * math.log(2 * math.pi))
    assert orc.energy(k, 0.3, X, y) == pytest.approx(expect, rel=1e-12)


def test_dirac_fires_on_duplicates_and_uses_abs():
    # SURVEY Q4 (DiracKernel.scala:43-47)
    X = np.array([[0.0, 1.0], [0.0, 1.0], [2.0, 2.0]])
    K = orc.dirac(-0.7).value(X, X)
    assert np.array_equal(K, np.array([[0.7, 0.7, 0.0], [0.7, 0.7, 0.0], [0.0, 0.0, 0.7]]))


def test_mahalanobis_reference_gradient_is_the_defective_one():
    # SURVEY D2 (RBFKernel.scala:142-143): every bandwidth gets k*|x-y|^2/b_i^3
    X, _, _ = orc.synthetic(5, 2, seed=2)
    h = {"MahalanobisAmplitude": 1.0, "MahalanobisBandwidth_0": 1.5, "MahalanobisBandwidth_1": 0.5}
    g = orc.Kernel("ard", h, "reference").grads(X, X)
    assert np.allclose(g["MahalanobisBandwidth_0"] * 1.5 ** 3, g["MahalanobisBandwidth_1"] * 0.5 ** 3)
    ge = orc.Kernel("ard", h, "exact").grads(X, X)
    assert not np.allclose(ge["MahalanobisBandwidth_0"] * 1.5 ** 3, ge["MahalanobisBandwidth_1"] * 0.5 ** 3)


def test_golden_fixtures_are_current():
    """The committed fixtures are what the oracle produces today (guards against silent drift)."""
    z = np.load(GOLD / "housing.npz")
    k = orc.Kernel("se", {"bandwidth": float(z["bandwidth"]), "amplitude": float(z["amplitude"])})
    assert orc.energy(k, float(z["noise"]), z["Xtr"], z["ytr"]) == pytest.approx(float(z["energy"]), rel=1e-12)
    s = np.load(GOLD / "synthetic.npz")
    k2 = orc.Kernel("matern", {"p": 2.0, "l": 2.2})
    assert orc.energy(k2, float(s["noise"]), s["X"], s["y"]) == pytest.approx(float(s["matern2_energy"]), rel=1e-12)


def test_extended_fixture_is_current():
    """tests/golden/extended.npz (the SURVEY 8f rows) is what the oracle produces today."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    z = np.load(GOLD / "extended.npz")
    X, y, noise = z["X"], z["y"], float(z["noise"])
    for name, k in mg.extended_kernels().items():
        assert orc.energy(k, noise, X, y) == pytest.approx(float(z[f"{name}_energy"]), rel=1e-12)
    se = orc.Kernel("se", {"bandwidth": 1.6, "amplitude": 1.1})
    assert orc.esgp_energy(se, noise, X, y, 0.6, 0.3) == pytest.approx(float(z["esgp_energy"]), rel=1e-12)
    assert orc.kronecker_mogp_energy(orc.Kernel("matern", {"p": 2.0, "l": 1.7}), 0.4, orc.Kernel("cauchy", {"sigma": 2.0}),
                                     X, z["Y"]) == pytest.approx(float(z["mogp_energy"]), rel=1e-12)
    assert orc.stp_energy(4.5, se, noise, X, y) == pytest.approx(float(z["stp_energy"]), rel=1e-12)
    # lambda = 0 collapses the extended skew GP onto the GP (SkewGaussian.scala:218-242 with alpha = 0)
    assert orc.esgp_energy(se, noise, X, y, 0.0, 0.3) == pytest.approx(orc.energy(se, noise, X, y), rel=1e-13)


def test_non_pd_gives_inf_and_nan():
    X = np.array([[0.0], [0.0], [1.0]])  # duplicate point, no noise: singular K
    y = np.array([1.0, 2.0, 3.0])
    k = orc.Kernel("rbf", {"bandwidth": 1.0})
    assert orc.energy(k, 0.0, X, y) == float("inf")
    g = orc.grad_energy(k, 0.0, X, y)
    assert all(math.isnan(v) for v in g.values())


# ---- Student-t process sibling: the reference's STPSpec known answers ----------------
def test_stpspec_likelihood():
    # STPSpec.scala:13-39: Dirac(0.5)+Dirac(0.5), mean v(0), mu = 4 -> state 6 -> energy() evaluates at 8
    X = np.array([[0.0], [1.0]])
    y = np.array([0.0, 1.0])
    e = orc.stp_energy(8.0, orc.Kernel("dirac", {"noiseLevel": 0.5}), 0.5, X, y, mean=X[:, 0])
    assert abs(e - (-math.log(1 / (2.0 * math.pi)))) < 1e-5


def test_stpspec_posterior_and_error_bars():
    # STPSpec.scala:41-80: one training point, SE(1,1), noise 0, degrees of freedom 4 + 2
    k = orc.Kernel("se", {"bandwidth": 1.0, "amplitude": 1.0})
    mu, mean, cov = orc.stp_predictive(6.0, k, 0.0, np.array([[1.0]]), np.array([1.0]), np.array([[2.0]]))
    rho = math.exp(-0.5)
    assert mean[0] == rho
    assert abs(cov[0, 0] * mu / (mu - 2.0) - (1.0 - rho * rho) * 7 / 5) < 1e-5
    lo, hi = orc.stp_error_bars(mu, mean, cov, 1.0)
    assert lo[0] == pytest.approx(rho - math.sqrt((1.0 - rho * rho) * 7 / 5), abs=1e-15)
    assert hi[0] == pytest.approx(rho + math.sqrt((1.0 - rho * rho) * 7 / 5), abs=1e-15)
