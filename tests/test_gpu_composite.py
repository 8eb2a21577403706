"""Kernel algebra on the device: `+`, `*`, `* c` of LocalScalarKernel
(CORE/kernels/LocalScalarKernel.scala:53-91, AdditiveCovariance.scala, MultiplicativeCovariance.scala) as
DGP_KERNEL_COMPOSITE, against the oracle's Sum / Prod / Scaled through the same model API."""
import math

import numpy as np
import pytest

import dynaml_b200 as dg
from dynaml_b200 import _capi
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL_K, TOL = 1e-12, 1e-9


def entry_rel(a, b):
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1e-290)).max())


def expressions(d):
    """name -> (device kernel, oracle kernel); leaves in the same order on both sides."""
    band = np.linspace(1.5, 3.0, d)
    ard_h = {"MahalanobisAmplitude": 1.1, **{f"MahalanobisBandwidth_{i}": b for i, b in enumerate(band)}}
    se = lambda: (dg.SEKernel(1.7, 1.2), orc.Kernel("se", {"bandwidth": 1.7, "amplitude": 1.2}))  # noqa: E731
    rbf = lambda: (dg.RBFKernel(2.4), orc.Kernel("rbf", {"bandwidth": 2.4}))  # noqa: E731
    mat = lambda: (dg.GenericMaternKernel(2.2, 2), orc.Kernel("matern", {"p": 2.0, "l": 2.2}))  # noqa: E731
    cau = lambda: (dg.CauchyKernel(2.3), orc.Kernel("cauchy", {"sigma": 2.3}))  # noqa: E731
    lap = lambda: (dg.LaplacianKernel(3.1), orc.Kernel("laplacian", {"beta": 3.1}))  # noqa: E731
    rq = lambda: (dg.RationalQuadraticKernel(1.5, 1.9), orc.Kernel("ratquad", {"mu": 1.5, "l": 1.9}))  # noqa: E731
    ard = lambda mode: (dg.MahalanobisKernel(band, 1.1, _capi.GRAD_EXACT if mode == "exact" else _capi.GRAD_REFERENCE),  # noqa: E731
                        orc.Kernel("ard", ard_h, mode))
    out = {}
    (a, oa), (b, ob) = se(), mat()
    out["se+matern"] = (a + b, orc.Sum(oa, ob))
    (a, oa), (b, ob) = se(), cau()
    out["se*cauchy"] = (a * b, orc.Prod(oa, ob))
    (a, oa) = mat()
    out["matern*2.5"] = (a * 2.5, orc.Scaled(oa, 2.5))
    (a, oa), (b, ob), (c, oc) = se(), lap(), rbf()
    out["se*laplacian+rbf*0.5"] = (a * b + c * 0.5, orc.Sum(orc.Prod(oa, ob), orc.Scaled(oc, 0.5)))
    (a, oa), (b, ob), (c, oc) = rbf(), rq(), mat()
    out["(rbf+ratquad)*matern"] = ((a + b) * c, orc.Prod(orc.Sum(oa, ob), oc))
    (a, oa), (b, ob) = ard("reference"), mat()
    out["ard_ref+matern"] = (a + b, orc.Sum(oa, ob))
    (a, oa), (b, ob) = ard("exact"), se()
    out["ard_exact*se"] = (a * b, orc.Prod(oa, ob))
    (a, oa) = se()
    out["se*se"] = (a * a, None)  # the same instance twice: exponent 2 (checked against se^2 below)
    return out


NAMES = ["se+matern", "se*cauchy", "matern*2.5", "se*laplacian+rbf*0.5", "(rbf+ratquad)*matern", "ard_ref+matern",
         "ard_exact*se"]


@pytest.mark.parametrize("name", NAMES)
def test_composite_kernel_matrices(name):
    X, _, _ = orc.synthetic(300, 5, seed=11)
    Xs, _, _ = orc.synthetic(140, 5, seed=12)
    k, ok = expressions(5)[name]
    K = k.buildKernelMatrix(X, len(X)).getKernelMatrix()
    assert entry_rel(K, orc.build_kernel_matrix(ok, X)) <= TOL_K
    assert np.array_equal(K, K.T)
    C = k.buildCrossKernelMatrix(X, Xs)
    assert entry_rel(C, orc.cross_kernel_matrix(ok, X, Xs)) <= TOL_K


def test_same_kernel_twice_is_a_square():
    X, _, _ = orc.synthetic(150, 3, seed=5)
    k, _ = expressions(3)["se*se"]
    K = k.buildKernelMatrix(X, len(X)).getKernelMatrix()
    K1 = orc.build_kernel_matrix(orc.Kernel("se", {"bandwidth": 1.7, "amplitude": 1.2}), X)
    assert entry_rel(K, K1 * K1) <= TOL_K


@pytest.mark.parametrize("name", NAMES)
def test_composite_energy_gradient_and_posterior(name):
    n, d = 260, 4
    X, y, Xs = orc.synthetic(n, d, seed=21, m=40)
    k, ok = expressions(d)[name]
    noise = 0.3
    model = dg.GPRegression(k, dg.DiracKernel(noise), list(zip(X, y)))
    h = model._current_state
    assert model.energy(h) == pytest.approx(orc.energy(ok, noise, X, y), rel=TOL)
    g, go = model.gradEnergy(h), orc.grad_energy(ok, noise, X, y)
    # same order on both sides: effective leaves of the expression, then the noise level
    assert len(g) == len(go)
    scale = max(abs(v) for v in go.values())
    for (gn, gv), (on, ov) in zip(g.items(), go.items()):
        assert gn.split("/")[-1] == on.split("/")[-1]
        assert gv == pytest.approx(ov, rel=TOL, abs=TOL * scale), (gn, on)
    post = model.predictiveDistribution(Xs)
    mu, cov = orc.predictive(ok, noise, X, y, Xs)
    assert np.abs(post.mu - mu).max() <= TOL * max(1.0, np.abs(mu).max())
    assert np.abs(post.covariance - cov).max() <= TOL * np.abs(cov).max()


def test_composite_with_dirac_factor_and_duplicates():
    """A DiracKernel inside the expression keeps the reference's |x-y| == 0 rule (duplicates included)."""
    X, y, _ = orc.synthetic(200, 3, seed=8)
    X[50] = X[7]
    se, dr = dg.SEKernel(1.3, 1.0), dg.DiracKernel(0.4)
    k = se * 0.7 + dr
    ok = orc.Sum(orc.Scaled(orc.Kernel("se", {"bandwidth": 1.3, "amplitude": 1.0}), 0.7), orc.dirac(0.4))
    # `<scaled kernel> + Dirac` is not the single-kernel noise path: the Dirac term is a base kernel of the composite
    K = k.buildKernelMatrix(X, len(X)).getKernelMatrix()
    assert entry_rel(K, orc.build_kernel_matrix(ok, X)) <= TOL_K
    assert K[50, 7] == pytest.approx(0.7 + 0.4, rel=1e-15)
    X[50] += 0.25  # with the duplicate K is singular by the reference's own rule; the gradient needs distinct points
    model = dg.GPRegression(k, dg.DiracKernel(0.2), list(zip(X, y)))
    g, go = model.gradEnergy(model._current_state), orc.grad_energy(ok, 0.2, X, y)
    scale = max(abs(v) for v in go.values())
    for (gn, gv), (on, ov) in zip(g.items(), go.items()):
        assert gv == pytest.approx(ov, rel=TOL, abs=TOL * scale), (gn, on)


def test_composite_batch_and_grid_search(ctx):
    X, y, _ = orc.synthetic(300, 3, seed=31)
    se, per = dg.SEKernel(1.5, 1.0), dg.CauchyKernel(2.0)
    k = se + per
    model = dg.GPRegression(k, dg.DiracKernel(0.25), list(zip(X, y)))
    names = model._hyper_parameters
    cands = []
    for s in (0.8, 1.0, 1.3, 1.7):
        c = dict(model._current_state)
        c[[n for n in names if n.endswith("bandwidth")][0]] *= s
        cands.append(c)
    es = model.energyBatch(cands)
    for c, e in zip(cands, es):
        lsc = c[[n for n in names if n.endswith("bandwidth")][0]]
        ok = orc.Sum(orc.Kernel("se", {"bandwidth": lsc, "amplitude": 1.0}), orc.Kernel("cauchy", {"sigma": 2.0}))
        assert e == pytest.approx(orc.energy(ok, 0.25, X, y), rel=TOL)


def test_composite_limits_and_malformed_programs(ctx):
    ks = [dg.RBFKernel(1.0 + 0.1 * i) for i in range(33)]
    big = ks[0]
    for k in ks[1:]:
        big = big + k
    with pytest.raises(ValueError):
        big._spec()
    X, y, _ = orc.synthetic(40, 2, seed=1)
    data = ctx.data_create(X, y)
    try:
        for prog in ([1.0], [1.0, 1.0, 99.0, 0.0, 1.0, 1.0, 1.0, 1.0], [1.0, 1.0, 1.0, 0.0, 1.0, 2.0, 1.0],
                     [1.0, 1.0, 1.0, 0.0, 1.0, 2.0, 1.0, 1.5], [1.0, 1.0, 10.0, 0.0, 0.0, 1.0, 1.0]):
            spec = _capi.make_spec(_capi.KERNEL_COMPOSITE, prog, 0.1)
            with pytest.raises(dg.DgpError):
                ctx.energy(data, spec)
    finally:
        ctx.data_destroy(data)


def test_composite_many_features_and_ragged_n():
    """d = 40 (> 32 features, generic loops) and N not a multiple of the tile."""
    X, y, _ = orc.synthetic(333, 40, seed=2)
    k = dg.SEKernel(6.0, 1.0) * dg.GenericMaternKernel(9.0, 1) + dg.LaplacianKernel(30.0) * 0.3
    ok = orc.Sum(orc.Prod(orc.Kernel("se", {"bandwidth": 6.0, "amplitude": 1.0}), orc.Kernel("matern", {"p": 1.0, "l": 9.0})),
                 orc.Scaled(orc.Kernel("laplacian", {"beta": 30.0}), 0.3))
    model = dg.GPRegression(k, dg.DiracKernel(0.2), list(zip(X, y)))
    assert model.energy(model._current_state) == pytest.approx(orc.energy(ok, 0.2, X, y), rel=TOL)
    g, go = model.gradEnergy(model._current_state), orc.grad_energy(ok, 0.2, X, y)
    scale = max(abs(v) for v in go.values())
    for (gn, gv), (on, ov) in zip(g.items(), go.items()):
        assert gv == pytest.approx(ov, rel=TOL, abs=TOL * scale), (gn, on)


# ---- the two non-stationary kernels (PolynomialKernel, FBMKernel): evaluated from the uncentred points -----------
def test_polynomial_and_fbm_matrices_on_uncentred_points():
    rng = np.random.default_rng(41)
    X = rng.standard_normal((300, 4)) * 0.6 + 1.5   # far from the origin: centring must not leak in
    Xs = rng.standard_normal((130, 4)) * 0.6 + 1.5
    for k, ok in ((dg.PolynomialKernel(3, 1.5), orc.Kernel("polynomial", {"degree": 3.0, "offset": 1.5})),
                  (dg.FBMKernel(0.7), orc.Kernel("fbm", {"hurst": 0.7}))):
        K = k.buildKernelMatrix(X, len(X)).getKernelMatrix()
        assert entry_rel(K, orc.build_kernel_matrix(ok, X)) <= TOL_K
        C = k.buildCrossKernelMatrix(X, Xs)
        Co = orc.cross_kernel_matrix(ok, X, Xs)
        assert float(np.abs(C - Co).max() / np.abs(Co).max()) <= TOL_K  # FBM cross entries cancel towards 0
    assert dg.PolynomialKernel(2, 1.0).evaluate([1.0, 2.0], [0.5, -1.0]) == pytest.approx((0.5 - 2.0 + 1.0) ** 2, rel=1e-15)


def test_polynomial_model_energy_gradient_and_posterior():
    X, y, Xs = orc.synthetic(220, 3, seed=43, m=30)
    X, Xs = X + 0.8, Xs + 0.8
    k, ok = dg.PolynomialKernel(2, 1.0), orc.Kernel("polynomial", {"degree": 2.0, "offset": 1.0})
    model = dg.GPRegression(k, dg.DiracKernel(0.5), list(zip(X, y)))
    h = model._current_state
    assert model.energy(h) == pytest.approx(orc.energy(ok, 0.5, X, y), rel=TOL)
    g, go = model.gradEnergy(h), orc.grad_energy(ok, 0.5, X, y)
    assert g["degree"] == 0.0 and g["offset"] == 0.0          # the reference's stub gradient
    assert g["noiseLevel"] == pytest.approx(go["noiseLevel"], rel=TOL)
    post = model.predictiveDistribution(Xs)
    mu, cov = orc.predictive(ok, 0.5, X, y, Xs)
    assert np.abs(post.mu - mu).max() <= TOL * max(1.0, np.abs(mu).max())
    assert np.abs(post.covariance - cov).max() <= TOL * np.abs(cov).max()


def test_fbm_gradient_reference_is_nan_and_exact_matches_oracle():
    X, y, _ = orc.synthetic(200, 2, seed=47)
    X = X + 2.0
    for mode, cmode in (("reference", _capi.GRAD_REFERENCE), ("exact", _capi.GRAD_EXACT)):
        k, ok = dg.FBMKernel(0.6, cmode), orc.Kernel("fbm", {"hurst": 0.6}, mode)
        model = dg.GPRegression(k, dg.DiracKernel(0.4), list(zip(X, y)))
        h = model._current_state
        assert model.energy(h) == pytest.approx(orc.energy(ok, 0.4, X, y), rel=TOL)
        g, go = model.gradEnergy(h), orc.grad_energy(ok, 0.4, X, y)
        if mode == "reference":
            assert math.isnan(g["hurst"]) and math.isnan(go["hurst"])   # 0^H ln 0 on the diagonal
        else:
            assert g["hurst"] == pytest.approx(go["hurst"], rel=TOL)
        assert g["noiseLevel"] == pytest.approx(go["noiseLevel"], rel=TOL)


def test_polynomial_times_se_composite():
    X, y, _ = orc.synthetic(200, 3, seed=49)
    X = X + 0.5
    k = dg.PolynomialKernel(2, 1.0) * dg.SEKernel(1.5, 1.0)
    ok = orc.Prod(orc.Kernel("polynomial", {"degree": 2.0, "offset": 1.0}),
                  orc.Kernel("se", {"bandwidth": 1.5, "amplitude": 1.0}))
    model = dg.GPRegression(k, dg.DiracKernel(0.3), list(zip(X, y)))
    assert model.energy(model._current_state) == pytest.approx(orc.energy(ok, 0.3, X, y), rel=TOL)
    g, go = model.gradEnergy(model._current_state), orc.grad_energy(ok, 0.3, X, y)
    scale = max(abs(v) for v in go.values())
    for (gn, gv), (on, ov) in zip(g.items(), go.items()):
        assert gv == pytest.approx(ov, rel=TOL, abs=TOL * scale), (gn, on)


def test_general_noise_kernel_enters_training_matrix_only():
    """AbstractGPRegressionModel takes any LocalScalarKernel as noise model (:269-295): K = cov + noise on the
    training set, the covariance alone for K* and K**; gradients over both kernels' hyper-parameters."""
    X, y, Xs = orc.synthetic(230, 3, seed=91, m=35)
    cov, ocov = dg.SEKernel(1.6, 1.1), orc.Kernel("se", {"bandwidth": 1.6, "amplitude": 1.1})
    noise = dg.LaplacianKernel(0.4) * 0.2 + dg.DiracKernel(0.3)          # short-range correlated + white noise
    onoise = orc.Sum(orc.Scaled(orc.Kernel("laplacian", {"beta": 0.4}), 0.2), orc.dirac(0.3))
    model = dg.GPRegression(cov, noise, list(zip(X, y)))
    K = orc.build_kernel_matrix(ocov, X) + orc.build_kernel_matrix(onoise, X)
    h = model._current_state
    assert model.energy(h) == pytest.approx(orc.log_likelihood(y, K), rel=TOL)
    # posterior: K*, K** without the noise kernel
    import scipy.linalg as sla
    L = sla.cholesky(K, lower=True)
    Ks, Kss = orc.cross_kernel_matrix(ocov, X, Xs), orc.build_kernel_matrix(ocov, Xs)
    alpha = sla.cho_solve((L, True), y)
    v = sla.solve_triangular(L, Ks, lower=True)
    post = model.predictiveDistribution(Xs)
    assert np.abs(post.mu - Ks.T @ alpha).max() <= TOL * np.abs(Ks.T @ alpha).max()
    assert np.abs(post.covariance - (Kss - v.T @ v)).max() <= TOL * np.abs(Kss).max()
    # gradient over covariance + noiseModel
    g = model.gradEnergy(h)
    go = orc.grad_energy(orc.Sum(ocov, onoise), 0.0, X, y, noise_blocked=True)
    ref = list(go.values())   # bandwidth, amplitude, beta, noiseLevel (of the Dirac inside the noise model)
    got = list(g.values())
    assert len(got) == 4
    scale = max(abs(v_) for v_ in ref)
    assert np.allclose(got, ref, rtol=TOL, atol=TOL * scale)
    # persisted model keeps the cross spec
    model.persist(h)
    assert model.predict(Xs[0]) == pytest.approx(float((Ks.T @ alpha)[0]), rel=TOL)


# ---- kernels on tuple inputs: TensorCombinationKernel / KroneckerProductKernel over concatenated vectors ---------
def test_tensor_combination_kernels():
    X, y, Xs = orc.synthetic(240, 5, seed=93, m=30)
    band = [1.5, 2.5, 2.0]
    ard_h = {"MahalanobisAmplitude": 1.1, **{f"MahalanobisBandwidth_{i}": b for i, b in enumerate(band)}}
    cases = {
        "se (x) laplacian": (dg.KroneckerProductKernel(dg.SEKernel(1.4, 1.0), dg.LaplacianKernel(2.0), 2),
                             orc.Prod(orc.Sliced(orc.Kernel("se", {"bandwidth": 1.4, "amplitude": 1.0}), 0, 2),
                                      orc.Sliced(orc.Kernel("laplacian", {"beta": 2.0}), 2))),
        "matern (+) ard_exact": (dg.TensorCombinationKernel(dg.GenericMaternKernel(1.8, 1),
                                                            dg.MahalanobisKernel(band, 1.1, _capi.GRAD_EXACT), 2, "+"),
                                 orc.Sum(orc.Sliced(orc.Kernel("matern", {"p": 1.0, "l": 1.8}), 0, 2),
                                         orc.Sliced(orc.Kernel("ard", ard_h, "exact"), 2))),
        "(rbf (x) cauchy) (x) rbf": (dg.KroneckerProductKernel(
            dg.KroneckerProductKernel(dg.RBFKernel(1.6), dg.CauchyKernel(2.2), 2), dg.RBFKernel(2.4), 4),
            orc.Prod(orc.Prod(orc.Sliced(orc.Kernel("rbf", {"bandwidth": 1.6}), 0, 2),
                              orc.Sliced(orc.Kernel("cauchy", {"sigma": 2.2}), 2, 2)),
                     orc.Sliced(orc.Kernel("rbf", {"bandwidth": 2.4}), 4))),
    }
    for name, (k, ok) in cases.items():
        K = k.buildKernelMatrix(X, len(X)).getKernelMatrix()
        assert entry_rel(K, orc.build_kernel_matrix(ok, X)) <= TOL_K, name
        model = dg.GPRegression(k, dg.DiracKernel(0.3), list(zip(X, y)))
        h = model._current_state
        assert model.energy(h) == pytest.approx(orc.energy(ok, 0.3, X, y), rel=TOL), name
        g, go = model.gradEnergy(h), orc.grad_energy(ok, 0.3, X, y)
        assert len(g) == len(go), name
        scale = max(abs(v) for v in go.values())
        for (gn, gv), (on, ov) in zip(g.items(), go.items()):
            assert gv == pytest.approx(ov, rel=TOL, abs=TOL * scale), (name, gn, on)
        post = model.predictiveDistribution(Xs)
        mu, cov = orc.predictive(ok, 0.3, X, y, Xs)
        assert np.abs(post.mu - mu).max() <= TOL * max(1.0, np.abs(mu).max()), name
        assert np.abs(post.covariance - cov).max() <= TOL * np.abs(cov).max(), name


def test_tensor_kernel_range_errors(ctx):
    X, y, _ = orc.synthetic(40, 3, seed=1)
    data = ctx.data_create(X, y)
    try:
        bad = dg.KroneckerProductKernel(dg.RBFKernel(1.0), dg.RBFKernel(1.0), 5)   # split beyond d = 3
        with pytest.raises(dg.DgpError):
            ctx.energy(data, bad._spec(noise=0.1))
        poly = dg.KroneckerProductKernel(dg.PolynomialKernel(2, 1.0), dg.RBFKernel(1.0), 2)
        with pytest.raises(dg.DgpError):
            ctx.energy(data, poly._spec(noise=0.1))                                # inner-product kernel on a slice
    finally:
        ctx.data_destroy(data)


def test_general_noise_with_ard_covariance_cross_spec():
    """The covariance-only cross spec of an ARD-SE kernel keeps its own pre-scaled copy of the training points."""
    import scipy.linalg as sla
    X, y, Xs = orc.synthetic(200, 3, seed=95, m=25)
    band = [1.5, 2.5, 2.0]
    cov = dg.MahalanobisKernel(band, 1.2)
    ocov = orc.Kernel("ard", {"MahalanobisAmplitude": 1.2, **{f"MahalanobisBandwidth_{i}": b for i, b in enumerate(band)}})
    noise = dg.CauchyKernel(0.3) * 0.15 + dg.DiracKernel(0.25)
    onoise = orc.Sum(orc.Scaled(orc.Kernel("cauchy", {"sigma": 0.3}), 0.15), orc.dirac(0.25))
    model = dg.GPRegression(cov, noise, list(zip(X, y)))
    K = orc.build_kernel_matrix(ocov, X) + orc.build_kernel_matrix(onoise, X)
    assert model.energy(model._current_state) == pytest.approx(orc.log_likelihood(y, K), rel=TOL)
    L = sla.cholesky(K, lower=True)
    Ks, Kss = orc.cross_kernel_matrix(ocov, X, Xs), orc.build_kernel_matrix(ocov, Xs)
    alpha = sla.cho_solve((L, True), y)
    v = sla.solve_triangular(L, Ks, lower=True)
    post = model.predictiveDistribution(Xs)
    assert np.abs(post.mu - Ks.T @ alpha).max() <= TOL * np.abs(Ks.T @ alpha).max()
    assert np.abs(post.covariance - (Kss - v.T @ v)).max() <= TOL * np.abs(Kss).max()


def test_cov_plus_dirac_as_model_covariance_with_zero_model_noise():
    """`k + DiracKernel` used as the covariance of a model whose own noise is DiracKernel(0.0): the sum has to stay a
    composite inside the model so that the gradient lines up with the covariance's hyper-parameter list."""
    X, y, _ = orc.synthetic(200, 3, seed=4)
    cov = dg.SEKernel(1.4, 1.1) + dg.DiracKernel(0.3)
    model = dg.GPRegression(cov, dg.DiracKernel(0.0), list(zip(X, y)))
    ok = orc.Kernel("se", {"bandwidth": 1.4, "amplitude": 1.1})
    assert model.energy(model._current_state) == pytest.approx(orc.energy(ok, 0.3, X, y), rel=TOL)
    g, go = model.gradEnergy(model._current_state), orc.grad_energy(ok, 0.3, X, y)
    scale = max(abs(v) for v in go.values())
    # keys keep the leaf prefix of the sum ("<SEKernel@..>/bandwidth"), as the reference's h.split("/").tail does (:239)
    by_leaf = {k.split("/")[-1]: v for k, v in g.items() if not k.endswith("noiseLevel")}
    for name in ("bandwidth", "amplitude"):
        assert by_leaf[name] == pytest.approx(go[name], rel=TOL, abs=TOL * scale)
    assert any(k.endswith("noiseLevel") for k in g)
    # the validation-set copy of the data is released with the model's other device blocks
    model.validationSet_(list(zip(X[:20], y[:20])))
    model.persist(model._current_state)
    assert model._train_data is not None
    model._drop_data()
    assert model._train_data is None and model._data is None
