"""Context housekeeping through the C ABI: dgp_ctx_info, dgp_ctx_trim (handles stay valid, results unchanged),
dgp_dist_info on an emulated grid, and the launch counter the bench reports as gpu_launches."""
import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    c = dg.Context(0)
    yield c
    c.close()


def test_info_reports_the_device(ctx):
    info = ctx.info()
    assert info["device"] == 0
    assert info["sm_count"] >= 100          # B200: 148
    assert 0 < info["mem_free"] <= info["mem_total"]
    assert info["mem_total"] > 100 * 2 ** 30


def test_trim_keeps_handles_and_results(ctx):
    X, y, Xs = orc.synthetic(1200, 4, seed=3, m=50)
    k = dg.SEKernel(1.7, 1.1)
    spec = k._spec(noise=0.2)
    data = ctx.data_create(X, y)
    try:
        e0, g0 = ctx.energy_grad(data, spec)
        model = ctx.fit(data, spec)
        mu0, cov0, _, _ = ctx.predict(model, Xs)
        free_before = ctx.info()["mem_free"]
        ctx.trim()
        assert ctx.info()["mem_free"] >= free_before       # the arena went back to the driver
        e1, g1 = ctx.energy_grad(data, spec)                # data handle still valid, workspaces re-grow
        mu1, cov1, _, _ = ctx.predict(model, Xs)            # model handle still valid
        assert e1 == e0
        assert np.array_equal(np.asarray(g1), np.asarray(g0))
        assert np.array_equal(mu1, mu0) and np.array_equal(cov1, cov0)
        ctx.model_destroy(model)
    finally:
        ctx.data_destroy(data)


def test_launch_counter_counts_kernels(ctx):
    X, y, _ = orc.synthetic(700, 3, seed=1)
    data = ctx.data_create(X, y)
    try:
        spec = dg.SEKernel(1.5, 1.0)._spec(noise=0.3)
        ctx.energy(data, spec)
        a = ctx.launch_count()
        ctx.energy(data, spec)
        b = ctx.launch_count()
        ctx.energy_grad(data, spec)
        c = ctx.launch_count()
        assert b - a >= 3                # assembly, at least one leaf, the solves
        assert c - b > b - a             # the gradient adds the inverse and the trace pass
    finally:
        ctx.data_destroy(data)


def test_dist_info_on_an_emulated_grid(ctx):
    ctx.dist_init(None, 0, 1, 2, 2)
    try:
        info = ctx.dist_info()
        assert info["grid"] == [2, 2]
        assert info["emulated"] is True
        assert info["nccl_ranks"] in (0, 1) and info["nccl_rank"] in (-1, 0)
    finally:
        ctx.dist_finalize()


def test_dist_info_without_init_is_an_error(ctx):
    with pytest.raises(dg.DgpError):
        ctx.dist_info()


def test_large_results_leave_through_the_staged_copy(ctx):
    """Results of 32 MB and more are copied out through pinned staging chunks on a stream of their own (api.cu,
    copy_out_2d): same values as the oracle, ragged sizes so that the last chunk is partial."""
    k, ok = dg.SEKernel(1.9, 1.2), orc.Kernel("se", {"bandwidth": 1.9, "amplitude": 1.2})
    X, y, Xs = orc.synthetic(2300, 3, seed=4, m=2101)
    C = k.buildCrossKernelMatrix(X, Xs)                      # 2300 x 2101 doubles = 38.7 MB
    Co = ok.value(X, Xs)
    assert C.shape == Co.shape
    assert float((np.abs(C - Co) / np.maximum(np.abs(Co), 1e-290)).max()) <= 1e-12
    model = dg.GPRegression(k, dg.DiracKernel(0.3), list(zip(X[:400], y[:400])))
    post = model.predictiveDistribution(Xs)                  # 2101 x 2101 covariance = 35.3 MB
    mu, cv = orc.predictive(ok, 0.3, X[:400], y[:400], Xs)
    assert float(np.abs(post.mu - mu).max()) <= 1e-9 * float(np.abs(mu).max())
    assert float(np.abs(np.asarray(post.covariance) - cv).max()) <= 1e-9 * float(np.abs(cv).max())
    res = model.predictionWithErrorBars(list(Xs[:40]), 2)    # bars enqueued before a (small) copy-out
    lo_o, hi_o = orc.error_bars(*orc.predictive(ok, 0.3, X[:400], y[:400], Xs[:40]), 2.0)
    assert np.allclose([r[2] for r in res], lo_o, rtol=0, atol=1e-8 * float(np.abs(hi_o).max()))
    model._drop_data()


def test_repeated_calls_are_bitwise_reproducible(ctx):
    """Race detection by determinism: every reduction in the library has a fixed order, so repeated evaluations must
    agree to the last bit -- whatever the interleaving of the look-ahead streams (Cholesky panel stream, solves under
    the inverse, triangular solve of the prediction) and of the staged copy-out."""
    X, y, Xs = orc.synthetic(4301, 6, seed=21, m=520)
    spec = dg.GenericMaternKernel(2.4, 2)._spec(noise=0.2)
    data = ctx.data_create(X, y)
    try:
        e0, g0 = ctx.energy_grad(data, spec)
        for _ in range(6):
            e, g = ctx.energy_grad(data, spec)
            assert e == e0 and np.array_equal(np.asarray(g), np.asarray(g0), equal_nan=True)
        eb, _ = ctx.energy_batch(data, [spec] * 7)
        assert all(v == e0 for v in eb)                       # concurrent slots: same bits as the single evaluation
        model = ctx.fit(data, spec)
        mu0, cov0, lo0, hi0 = ctx.predict(model, Xs, want_cov=True, want_bars=True, sigma=2.0)
        for _ in range(3):
            mu, cov, lo, hi = ctx.predict(model, Xs, want_cov=True, want_bars=True, sigma=2.0)
            assert np.array_equal(mu, mu0) and np.array_equal(cov, cov0)
            assert np.array_equal(lo, lo0) and np.array_equal(hi, hi0)
        ctx.model_destroy(model)
    finally:
        ctx.data_destroy(data)
