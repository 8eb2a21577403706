"""Parity of the CUDA building blocks (DMMA GEMM, blocked Cholesky, inverse) against NumPy/LAPACK,
through the C ABI.  FP64; tolerances are stated per test."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("shape", [(128, 128, 16), (200, 300, 150), (513, 129, 1000), (1, 1, 1)])
def test_gemm_layouts(ctx, ta, tb, shape):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    C, _ = ctx.test_gemm(A, B, C0, alpha=0.7, beta=-0.3, transA=ta, transB=tb)
    ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 0.3 * C0
    # K-term dot products in FP64: |err| <= ~K * eps * |a||b|
    assert np.abs(C - ref).max() <= 64 * K * np.finfo(float).eps * max(1.0, np.abs(ref).max())


def test_gemm_lower_only_leaves_upper_tiles_untouched(ctx):
    rng = np.random.default_rng(1)
    n, k = 384, 256
    A = rng.standard_normal((n, k))
    C0 = rng.standard_normal((n, n))
    C, _ = ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    ref = C0 - A @ A.T
    for ti in range(3):
        for tj in range(3):
            blk = (slice(ti * 128, ti * 128 + 128), slice(tj * 128, tj * 128 + 128))
            if ti >= tj:
                assert np.abs(C[blk] - ref[blk]).max() < 1e-11
            else:
                assert np.array_equal(C[blk], C0[blk])


def test_gemm_linearity_at_size(ctx):
    """Size-independent property at a bench-size K: (A+B)C = AC + BC to rounding."""
    rng = np.random.default_rng(2)
    n, k = 1024, 4096
    A, B, Cm = rng.standard_normal((n, k)), rng.standard_normal((n, k)), rng.standard_normal((k, 256))
    s, _ = ctx.test_gemm(A + B, Cm)
    a, _ = ctx.test_gemm(A, Cm)
    b, _ = ctx.test_gemm(B, Cm)
    assert np.abs(s - (a + b)).max() <= 1e-10 * np.abs(s).max()


def _spd(n, seed, jitter=1e-3):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n + 8))
    return G @ G.T + n * jitter * np.eye(n)


@pytest.mark.parametrize("n", [1, 5, 128, 129, 300, 512, 1000, 2048])
@pytest.mark.parametrize("lookahead", [0, 1])
def test_potrf_matches_lapack(ctx, n, lookahead):
    S = _spd(n, n)
    ctx.set_option("lookahead", lookahead)
    L, Linv, Ainv, logdet, _ = ctx.test_potrf(S, want_linv=True, want_ainv=True)
    ctx.set_option("lookahead", 1)
    Lr = sla.cholesky(S, lower=True)
    assert np.abs(L - Lr).max() <= 1e-12 * np.abs(Lr).max()
    assert logdet == pytest.approx(np.log(np.diag(Lr)).sum(), rel=1e-12, abs=1e-12)
    assert np.abs(Linv @ Lr - np.eye(n)).max() < 1e-10
    assert np.abs(Ainv @ S - np.eye(n)).max() < 1e-9
    assert np.array_equal(Ainv, Ainv.T)  # mirrored, bitwise symmetric


@pytest.mark.parametrize("nb", [128, 256, 512, 768])
def test_potrf_outer_block_sizes_agree(ctx, nb):
    S = _spd(1536, 9)
    ctx.set_option("nb", nb)
    L, _, _, logdet, _ = ctx.test_potrf(S)
    ctx.set_option("nb", 0)
    Lr = sla.cholesky(S, lower=True)
    assert np.abs(L - Lr).max() <= 1e-12 * np.abs(Lr).max()


def test_potrf_roundtrip_at_size(ctx):
    """L L' reproduces A at a size the CPU would not check entry-wise in the test budget."""
    n = 4096
    S = _spd(n, 4)
    L, _, _, _, _ = ctx.test_potrf(S)
    R, _ = ctx.test_gemm(L, L, transB=True)
    assert np.abs(R - S).max() <= 1e-12 * np.abs(S).max()


def test_potrf_reports_non_pd(ctx):
    import dynaml_b200 as dg
    S = _spd(300, 5)
    S[200, 200] = -1.0
    with pytest.raises(dg.NotPositiveDefinite):
        ctx.test_potrf(S)
    S2 = _spd(64, 6)
    S2[10, 10] = np.nan
    with pytest.raises(dg.NotPositiveDefinite):
        ctx.test_potrf(S2)


def test_measured_peaks_are_sane(ctx):
    p = ctx.measure_peaks()
    assert 10.0 < p["dmma_tflops"] < 80.0
    assert 2000.0 < p["hbm_write_gbs"] < 9000.0
