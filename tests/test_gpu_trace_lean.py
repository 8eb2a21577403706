"""The lean Gram trace pass of gradEnergy (grad.cu, grad_trace_lean_kernel) against the oracle and against the
exact-difference pass (option grad_lean = 0) at sizes where it is selected (padded N >= 4096): SE, ARD-SE in the
reference's gradient mode (D2), Matern orders 1-3; ragged N (the last tile row goes through the edge pass), duplicate
points and the exact ARD mode (both must fall back to exact differences).
Reference: AbstractGPRegressionModel.scala:231-241 (gradEnergy), GenericMaternKernel.scala:70-97, RBFKernel.scala:45-52,
134-145."""
import math

import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import cpu_baseline as cb
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    c = dg.default_context()
    yield c
    c.set_option("grad_lean", 1)
    c.trim()


def _cases():
    d = 5
    bw = [1.5 + 0.2 * i for i in range(d)]
    return {
        "se": (dg.SEKernel(2.0, 1.3), orc.Kernel("se", {"bandwidth": 2.0, "amplitude": 1.3}), d),
        "ard_ref": (dg.MahalanobisKernel(bw, 1.2),
                    orc.Kernel("ard", {"MahalanobisAmplitude": 1.2, **{f"MahalanobisBandwidth_{i}": bw[i] for i in range(d)}}), d),
        "matern1": (dg.GenericMaternKernel(1.8, 1), orc.Kernel("matern", {"p": 1.0, "l": 1.8}), 3),
        "matern2": (dg.GenericMaternKernel(2.5, 2), orc.Kernel("matern", {"p": 2.0, "l": 2.5}), 12),
        "matern3": (dg.GenericMaternKernel(2.2, 3), orc.Kernel("matern", {"p": 3.0, "l": 2.2}), 7),
    }


def _grad_by_name(kernel, g):
    names = list(kernel.hyper_parameters) + ["noiseLevel"]
    return dict(zip(names, g))


@pytest.mark.parametrize("name", ["se", "ard_ref", "matern1", "matern2", "matern3"])
@pytest.mark.parametrize("n", [4096, 4301])
def test_lean_trace_matches_oracle_and_exact_pass(ctx, name, n):
    kernel, ok, d = _cases()[name]
    X, y, _ = orc.synthetic(n, d, seed=n + d)
    noise = 0.3
    data = ctx.data_create(X, y)
    try:
        spec = kernel._spec(noise=noise)
        ctx.set_option("grad_lean", 0)
        e0, g0 = ctx.energy_grad(data, spec)
        launches = ctx.launch_count()
        ctx.energy_grad(data, spec)
        exact_launches = ctx.launch_count() - launches
        ctx.set_option("grad_lean", 1)
        e1, g1 = ctx.energy_grad(data, spec)
        launches = ctx.launch_count()
        ctx.energy_grad(data, spec)
        assert ctx.launch_count() - launches == exact_launches + 1    # lean pass + edge pass instead of one pass
        assert e1 == e0
        g0, g1 = np.asarray(g0), np.asarray(g1)
        fin = np.isfinite(g0)
        assert np.array_equal(fin, np.isfinite(g1))
        assert np.allclose(g1[fin], g0[fin], rtol=1e-11, atol=0)
        eo, go = cb.energy_grad_fast(ok, noise, X, y)
        assert e1 == pytest.approx(eo, rel=TOL)
        got = _grad_by_name(kernel, g1)
        for k, v in go.items():
            key = [q for q in got if q == k or q.endswith(k)]
            assert key, (k, list(got))
            assert got[key[0]] == pytest.approx(v, rel=TOL), k
    finally:
        ctx.set_option("grad_lean", 1)
        ctx.data_destroy(data)


def test_duplicate_points_fall_back_to_exact_differences(ctx):
    # a duplicated point pairs with its twin at distance zero off the diagonal: the noise moment must see it, so the
    # lean pass (which assumes zero distances on the diagonal only) may not be selected.
    n, d = 4200, 4
    X, y, _ = orc.synthetic(n, d, seed=11)
    X[1234] = X[17]
    kernel, ok = dg.SEKernel(1.9, 1.1), orc.Kernel("se", {"bandwidth": 1.9, "amplitude": 1.1})
    data = ctx.data_create(X, y)
    try:
        spec = kernel._spec(noise=0.4)
        ctx.set_option("grad_lean", 0)
        launches = ctx.launch_count()
        e0, g0 = ctx.energy_grad(data, spec)
        n0 = ctx.launch_count() - launches
        ctx.set_option("grad_lean", 1)
        launches = ctx.launch_count()
        e1, g1 = ctx.energy_grad(data, spec)
        assert ctx.launch_count() - launches == n0          # same kernels: the lean pass was not selected
        assert e1 == e0 and np.array_equal(np.asarray(g1), np.asarray(g0), equal_nan=True)
        # (Q4: the noise lands on the twin pair too, rows 17 and 1234 of K coincide and the matrix is singular, so
        # both passes answer +Inf / NaN or rounding noise; only the selection is under test here)
    finally:
        ctx.data_destroy(data)


def test_exact_ard_mode_keeps_the_exact_difference_pass(ctx):
    n, d = 4096, 4
    X, y, _ = orc.synthetic(n, d, seed=5)
    kernel = dg.MahalanobisKernel([1.5, 2.0, 2.5, 3.0], 1.0, grad_mode=dg._capi.GRAD_EXACT)
    data = ctx.data_create(X, y)
    try:
        spec = kernel._spec(noise=0.3)
        ctx.set_option("grad_lean", 0)
        e0, g0 = ctx.energy_grad(data, spec)
        ctx.set_option("grad_lean", 1)
        e1, g1 = ctx.energy_grad(data, spec)
        assert e1 == e0 and np.array_equal(np.asarray(g1), np.asarray(g0))
    finally:
        ctx.data_destroy(data)


def test_far_from_origin_inputs_fail_the_gram_gate_and_stay_exact(ctx):
    # |x|^2 ~ 1e8 relative to a unit bandwidth: gram_is_accurate() is false, assembly and trace use exact differences
    n, d = 4096, 3
    rng = np.random.default_rng(3)
    X = rng.standard_normal((n, d)) + np.array([1e4, -1e4, 1e4]) * rng.choice([-1.0, 1.0], size=(n, 1))
    y = np.sin(X[:, 0])
    kernel = dg.SEKernel(1.0, 1.0)
    data = ctx.data_create(X, y)
    try:
        spec = kernel._spec(noise=0.5)
        ctx.set_option("grad_lean", 0)
        e0, g0 = ctx.energy_grad(data, spec)
        ctx.set_option("grad_lean", 1)
        e1, g1 = ctx.energy_grad(data, spec)
        assert e1 == e0 and np.array_equal(np.asarray(g1), np.asarray(g0))
    finally:
        ctx.data_destroy(data)
