"""The two FP64 GEMM kernels behind every O(N^3) step -- the bulk-copy (TMA) / mbarrier kernel and the cp.async
kernel (dynaml_b200/csrc/gemm.cu) -- must agree BIT FOR BIT: both add the k-terms of an output element in the same
order, so any difference is a data hazard, not rounding.  Also the regression test of the stage hand-over hazard
(DESIGN.md 4.1): with the C-tile prefetch off, the earlier kernel returned a wrong factor in every N = 8192 run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DEFAULTS = {"gemm_variant": 4, "gemm_v4_mask": 3, "gemm_v4_bk32_mask": 2, "gemm_no_prefetch": 0}


@pytest.fixture
def gemm_options(ctx):
    def set_all(**kw):
        for k, v in {**DEFAULTS, **kw}.items():
            ctx.set_option(k, v)
    yield set_all
    set_all()


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("bk32", [0, 15])
def test_bulk_copy_kernel_equals_cp_async_kernel(ctx, gemm_options, ta, tb, bk32):
    rng = np.random.default_rng(7 + 2 * ta + tb)
    for (M, N, K) in ((128, 64, 32), (384, 320, 160), (640, 256, 1056), (200, 300, 150)):
        A = rng.standard_normal((K, M) if ta else (M, K))
        B = rng.standard_normal((N, K) if tb else (K, N))
        C0 = rng.standard_normal((M, N))
        # alpha, beta powers of two: both products of the epilogue's alpha * acc + beta * C are exact, so the result
        # does not depend on which of them the compiler contracts into the FMA
        gemm_options(gemm_variant=2)
        want, _ = ctx.test_gemm(A, B, C0, alpha=2.0, beta=-0.5, transA=ta, transB=tb)
        want0, _ = ctx.test_gemm(A, B, transA=ta, transB=tb)
        gemm_options(gemm_variant=4, gemm_v4_mask=15, gemm_v4_bk32_mask=bk32)
        got, _ = ctx.test_gemm(A, B, C0, alpha=2.0, beta=-0.5, transA=ta, transB=tb)
        got0, _ = ctx.test_gemm(A, B, transA=ta, transB=tb)
        assert np.array_equal(got, want) and np.array_equal(got0, want0), (M, N, K)
        ref = 2.0 * (A.T if ta else A) @ (B.T if tb else B) - 0.5 * C0
        assert np.abs(got - ref).max() <= 64 * K * np.finfo(float).eps * max(1.0, np.abs(ref).max())


def test_bulk_copy_kernel_lower_only_syrk(ctx, gemm_options):
    rng = np.random.default_rng(3)
    n, k = 1536, 768
    A = rng.standard_normal((n, k))
    C0 = rng.standard_normal((n, n))
    gemm_options(gemm_variant=2)
    want, _ = ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    gemm_options()
    got, _ = ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    assert np.array_equal(got, want)
    iu = np.triu_indices(n, 128)          # tiles strictly above the diagonal tiles are never written
    assert np.array_equal(got[iu], C0[iu])


def test_stage_handover_under_a_slow_epilogue(ctx, gemm_options):
    """N = 8192 energy (assembly + blocked Cholesky + solves): trailing updates fill every SM with two CTAs, and without
    the L2 prefetch of the C tile one CTA's read-modify-write epilogue backs up the load/store pipe while the other is
    in its main loop -- the condition under which a stage was refilled before its fragment loads had read it."""
    import dynaml_b200 as dg
    from oracle import gp_oracle as orc
    X, y, _ = orc.synthetic(8192, 8, seed=5)
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(3.0, 1.0)._spec(noise=0.05)
    try:
        gemm_options(gemm_variant=2)
        e_ref = ctx.energy(data, spec)
        assert np.isfinite(e_ref)
        for no_prefetch, reps in ((1, 10), (0, 4)):
            gemm_options(gemm_no_prefetch=no_prefetch)
            for _ in range(reps):
                assert ctx.energy(data, spec) == e_ref
    finally:
        ctx.data_destroy(data)


def test_energy_grad_identical_across_gemm_kernels(ctx, gemm_options):
    import dynaml_b200 as dg
    from oracle import gp_oracle as orc
    X, y, _ = orc.synthetic(3000, 6, seed=11)     # ragged: padded to 3072
    data = ctx.data_create(X, y)
    spec = dg.GenericMaternKernel(2.5, 2)._spec(noise=0.1)
    try:
        gemm_options(gemm_variant=2)
        e2, g2 = ctx.energy_grad(data, spec)
        gemm_options()
        e4, g4 = ctx.energy_grad(data, spec)
        assert e4 == e2
        assert all(a == b or (a != a and b != b) for a, b in zip(g4, g2))
        eo = orc.energy(orc.Kernel("matern", {"p": 2.0, "l": 2.5}), 0.1, X, y)
        assert abs(e4 - eo) <= 1e-9 * abs(eo)
    finally:
        ctx.data_destroy(data)


def test_transposed_inverse_equals_the_lower_one(ctx, gemm_options):
    """K^-1 for the gradient is U U^T with U = L^-T (NT / NN products only); the W = L^-1, W^T W route adds the same
    terms in the same order, so the two gradients agree bit for bit (option inverse_transposed)."""
    import dynaml_b200 as dg
    from oracle import gp_oracle as orc
    X, y, _ = orc.synthetic(2500, 7, seed=3)      # 2560 padded: a last, smaller join at every level
    data = ctx.data_create(X, y)
    spec = dg.MahalanobisKernel(np.linspace(1.5, 3.0, 7), 1.2)._spec(noise=0.05)
    try:
        ctx.set_option("inverse_transposed", 0)
        e0, g0 = ctx.energy_grad(data, spec)
        ctx.set_option("inverse_transposed", 1)
        e1, g1 = ctx.energy_grad(data, spec)
        assert e0 == e1 and np.array_equal(np.asarray(g0), np.asarray(g1))
        # and the factor's inverse itself against LAPACK
        rng = np.random.default_rng(0)
        A = rng.standard_normal((640, 640)); A = A @ A.T + 640 * np.eye(640)
        _, _, ainv, _, _ = ctx.test_potrf(A, want_ainv=True)
        assert np.abs(ainv - np.linalg.inv(A)).max() <= 1e-12 * np.abs(ainv).max()
    finally:
        ctx.set_option("inverse_transposed", 1)
        ctx.data_destroy(data)


def test_handover_modes_agree_and_the_plain_arrival_is_caught(ctx, gemm_options):
    """Option gemm_handover: the producer-side proxy fence (default), the consumer-side fence and the
    round-1 register fold return the cp.async kernel's result bit for bit; the plain arrival (mode 3) is the hazard
    itself -- if it ever stops failing, this test says so (the regression test above would then prove nothing)."""
    import dynaml_b200 as dg
    from oracle import gp_oracle as orc
    X, y, _ = orc.synthetic(8192, 8, seed=5)
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(3.0, 1.0)._spec(noise=0.05)
    try:
        gemm_options(gemm_variant=2)
        e_ref = ctx.energy(data, spec)
        gemm_options(gemm_no_prefetch=1)
        for mode in (1, 2, 0):
            ctx.set_option("gemm_handover", mode)
            assert all(ctx.energy(data, spec) == e_ref for _ in range(4)), mode
        ctx.set_option("gemm_handover", 3)
        wrong = sum(ctx.energy(data, spec) != e_ref for _ in range(4))
        assert wrong > 0, "the unordered hand-over no longer reproduces the hazard"
    finally:
        ctx.set_option("gemm_handover", 1)
        ctx.data_destroy(data)


def test_cholesky_and_solve_variants_agree(ctx):
    """Options of the mid-N chain: the chained triangular solves and the CUDA-graph replay of the factorisation are
    bit-identical to the step-by-step paths; the tile leaf and the scalar leaf agree to rounding; the staged bulk-store
    assembly writes the same entries."""
    import dynaml_b200 as dg
    from oracle import gp_oracle as orc
    X, y, _ = orc.synthetic(3000, 8, seed=2)            # ragged: padded to 3072
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(2.5, 1.1)._spec(noise=0.1)
    try:
        e0, g0 = ctx.energy_grad(data, spec)
        for key, val in (("trsv_chained", 0), ("graphs", 1), ("assembly_store", 1)):
            ctx.set_option(key, val)
            for _ in range(3):                           # graphs: eager, capture, replay
                e1, g1 = ctx.energy_grad(data, spec)
                assert e1 == e0 and np.array_equal(np.asarray(g1), np.asarray(g0)), (key, val)
            eb, st = ctx.energy_batch(data, [spec] * 5)
            assert all(v == ctx.energy(data, spec) for v in eb) and not st.any()
            ctx.set_option(key, 1 - val)
        ctx.set_option("leaf_variant", 0)
        e2, g2 = ctx.energy_grad(data, spec)
        assert e2 == pytest.approx(e0, rel=1e-13)
        assert np.allclose(np.asarray(g2), np.asarray(g0), rtol=1e-11, atol=1e-11 * np.abs(g0).max(), equal_nan=True)
        eo = orc.energy(orc.Kernel("se", {"bandwidth": 2.5, "amplitude": 1.1}), 0.1, X, y)
        assert abs(e2 - eo) <= 1e-9 * abs(eo) and abs(e0 - eo) <= 1e-9 * abs(eo)
    finally:
        for key, val in (("trsv_chained", 1), ("graphs", 0), ("assembly_store", 0), ("leaf_variant", 1)):
            ctx.set_option(key, val)
        ctx.data_destroy(data)
