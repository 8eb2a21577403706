"""The matrix builders of the SVMKernel companion object on the device:
buildPartitionedKernelMatrix / crossPartitonedKernelMatrix (SVMKernel.scala:181-248) block by block from resident
points, and buildKernelGradMatrix / buildCrossKernelGradMatrix (:107-179) with per-entry derivative parity."""
import math

import numpy as np
import pytest

import dynaml_b200 as dg
from dynaml_b200 import _capi
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL_K = 1e-12


def entry_rel(a, b):
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1e-290)).max())


def test_kernelspec_partitioned_builders_known_answers():
    """KernelSpec.scala:205-229: eval2(x, y) = 1 / (1 + (x - y)^2) -- the CauchyKernel with sigma = 1 -- on the points
    0, 1 with one element per block."""
    data = [[0.0], [1.0]]
    k = dg.CauchyKernel(1.0)
    k4 = dg.SVMKernel.buildPartitionedKernelMatrix(data, 2, 1, 1, k)
    assert k4.rows == 2 and k4.cols == 2 and k4.rowBlocks == 2 and k4.colBlocks == 2
    assert all(np.array_equal(b, [[1.0]] if i == j else [[0.5]]) for (i, j), b in k4._data)
    assert sorted(ij for ij, _ in k4._underlyingdata) == [(0, 0), (1, 0), (1, 1)]
    k5 = dg.SVMKernel.crossPartitonedKernelMatrix(data, [[0.0]], 1, 1, k)
    assert k5.rows == 2 and k5.cols == 1 and k5.rowBlocks == 2 and k5.colBlocks == 1
    assert all(np.array_equal(b, [[1.0]] if i == j else [[0.5]]) for (i, j), b in k5._data)
    k2 = dg.SVMKernel.buildSVMKernelMatrix(data, 2, k).getKernelMatrix()
    assert np.array_equal(k2, [[1.0, 0.5], [0.5, 1.0]])
    k3 = dg.SVMKernel.crossKernelMatrix(data, [[0.0]], k)
    assert k3.shape == (2, 1) and np.array_equal(k3[:, 0], [1.0, 0.5])


@pytest.mark.parametrize("name,n,d,block", [("se", 700, 5, 256), ("matern", 1300, 16, 1000), ("ard", 517, 8, 100),
                                            ("periodic", 300, 2, 77), ("sum", 400, 4, 150)])
def test_partitioned_blocks_match_the_oracle(name, n, d, block):
    X, _, Xs = orc.synthetic(n, d, seed=n, m=90)
    band = [1.5 + 0.2 * i for i in range(d)]
    pairs = {
        "se": (dg.SEKernel(1.9, 1.1), orc.Kernel("se", {"bandwidth": 1.9, "amplitude": 1.1})),
        "matern": (dg.GenericMaternKernel(3.0, 2), orc.Kernel("matern", {"p": 2.0, "l": 3.0})),
        "ard": (dg.MahalanobisKernel(band, 1.2),
                orc.Kernel("ard", {"MahalanobisAmplitude": 1.2, **{f"MahalanobisBandwidth_{i}": b for i, b in enumerate(band)}})),
        "periodic": (dg.PeriodicKernel(1.3, 1.4, 0.07), orc.Kernel("periodic", {"sigma": 1.3, "lengthscale": 1.4, "frequency": 0.07})),
        "sum": (dg.SEKernel(1.9, 1.1) + dg.CauchyKernel(2.0),
                orc.Sum(orc.Kernel("se", {"bandwidth": 1.9, "amplitude": 1.1}), orc.Kernel("cauchy", {"sigma": 2.0}))),
    }
    k, ok = pairs[name]
    noise = 0.25
    P = dg.SVMKernel.buildPartitionedKernelMatrix(X, n, block, block, k, noise=noise)
    blocks, rows, cols, rb, cb = orc.build_partitioned_kernel_matrix(ok, X, block, block, orc.dirac(noise))
    assert (P.rows, P.cols, P.rowBlocks, P.colBlocks) == (rows, cols, rb, cb)
    got = dict(P._underlyingdata)
    assert sorted(got) == sorted(blocks)
    for ij, b in blocks.items():
        assert got[ij].shape == b.shape and entry_rel(got[ij], b) <= TOL_K, ij
    for i in range(rb):                                   # diagonal blocks are bitwise symmetric (i >= j evaluated, mirrored)
        assert np.array_equal(got[(i, i)], got[(i, i)].T)
    assert entry_rel(P.toBreezeMatrix(), orc.build_kernel_matrix(ok, X, orc.dirac(noise))) <= TOL_K
    k.setBlockSizes((block, block))
    assert np.array_equal(k.buildBlockedKernelMatrix(X, n).toBreezeMatrix(),
                          dg.SVMKernel.buildPartitionedKernelMatrix(X, n, block, block, k).toBreezeMatrix())
    C = k.buildBlockedCrossKernelMatrix(X, Xs)
    cblocks, rows, cols, rb, cb = orc.cross_partitioned_kernel_matrix(ok, X, Xs, block, block)
    assert (C.rows, C.cols, C.rowBlocks, C.colBlocks) == (rows, cols, rb, cb)
    for ij, b in cblocks.items():
        assert entry_rel(dict(C._data)[ij], b) <= TOL_K


def test_partitioned_blocks_with_duplicated_points(ctx):
    """DiracKernel adds |noise| wherever |x - y| == 0: also in a block below the diagonal."""
    X, _, _ = orc.synthetic(300, 3, seed=2)
    X[250] = X[10]
    k, ok = dg.SEKernel(1.4, 1.0), orc.Kernel("se", {"bandwidth": 1.4, "amplitude": 1.0})
    P = dg.SVMKernel.buildPartitionedKernelMatrix(X, 300, 100, 100, k, noise=0.5)
    blocks, *_ = orc.build_partitioned_kernel_matrix(ok, X, 100, 100, orc.dirac(0.5))
    got = dict(P._underlyingdata)
    assert got[(2, 0)][50, 10] == pytest.approx(1.5, rel=1e-15)
    for ij, b in blocks.items():
        assert entry_rel(got[ij], b) <= TOL_K
    # the raw entry point: any block, not only the aligned ones of the partitioned builder
    h = ctx.data_create(X, np.zeros(300))
    try:
        K = orc.build_kernel_matrix(ok, X, orc.dirac(0.5))
        for (r0, nr, c0, nc) in [(0, 300, 0, 300), (37, 101, 5, 260), (200, 100, 0, 33), (290, 10, 290, 10), (3, 1, 299, 1)]:
            B = ctx.data_kernel_block(h, k._spec(noise=0.5), r0, nr, c0, nc)
            assert B.shape == (nr, nc) and entry_rel(B, K[r0:r0 + nr, c0:c0 + nc]) <= TOL_K, (r0, nr, c0, nc)
        with pytest.raises(_capi.DgpError):
            ctx.data_kernel_block(h, k._spec(noise=0.5), 250, 100, 0, 10)     # beyond the last point
    finally:
        ctx.data_destroy(h)


GRAD_CASES = {
    "se": (lambda d: dg.SEKernel(1.7, 1.2), lambda d: orc.Kernel("se", {"bandwidth": 1.7, "amplitude": 1.2})),
    "rbf": (lambda d: dg.RBFKernel(2.1), lambda d: orc.Kernel("rbf", {"bandwidth": 2.1})),
    "ard_ref": (lambda d: dg.MahalanobisKernel([1.5 + 0.3 * i for i in range(d)], 1.1),
                lambda d: orc.Kernel("ard", {"MahalanobisAmplitude": 1.1,
                                             **{f"MahalanobisBandwidth_{i}": 1.5 + 0.3 * i for i in range(d)}})),
    "ard_exact": (lambda d: dg.MahalanobisKernel([1.5 + 0.3 * i for i in range(d)], 1.1, _capi.GRAD_EXACT),
                  lambda d: orc.Kernel("ard", {"MahalanobisAmplitude": 1.1,
                                               **{f"MahalanobisBandwidth_{i}": 1.5 + 0.3 * i for i in range(d)}}, "exact")),
    "matern2": (lambda d: dg.GenericMaternKernel(2.2, 2), lambda d: orc.Kernel("matern", {"p": 2.0, "l": 2.2})),
    "periodic": (lambda d: dg.PeriodicKernel(1.3, 1.4, 0.07),
                 lambda d: orc.Kernel("periodic", {"sigma": 1.3, "lengthscale": 1.4, "frequency": 0.07})),
    "cauchy": (lambda d: dg.CauchyKernel(2.3), lambda d: orc.Kernel("cauchy", {"sigma": 2.3})),
    "laplacian": (lambda d: dg.LaplacianKernel(3.1), lambda d: orc.Kernel("laplacian", {"beta": 3.1})),
    "ratquad": (lambda d: dg.RationalQuadraticKernel(1.5, 1.9), lambda d: orc.Kernel("ratquad", {"mu": 1.5, "l": 1.9})),
}


@pytest.mark.parametrize("name", sorted(GRAD_CASES))
def test_gradient_matrices_entry_by_entry(name):
    """dK_ij / dtheta for every effective hyper-parameter against the oracle's Kernel.grads (the restatement of each
    kernel's gradientAt), <= 1e-12 of the largest entry; the kernel-matrix plane <= 1e-12 per entry."""
    n, d, m = 150, 5, 40
    X, _, Xs = orc.synthetic(n, d, seed=9, m=m)
    k, ok = GRAD_CASES[name][0](d), GRAD_CASES[name][1](d)
    hyp = k.effective_hyper_parameters
    got = dg.SVMKernel.buildKernelGradMatrix(X, hyp, k, noise=0.3)
    want = ok.grads(X, X)
    assert list(got) == ["kernel-matrix"] + hyp
    assert entry_rel(got["kernel-matrix"], orc.build_kernel_matrix(ok, X, orc.dirac(0.3))) <= TOL_K
    for h in hyp:
        scale = float(np.abs(want[h]).max()) or 1.0
        assert got[h].shape == (n, n) and float(np.abs(got[h] - want[h]).max()) <= TOL_K * scale, h
    full = dg.SVMKernel.buildKernelGradMatrix(X, hyp + ["noiseLevel"], k, noise=0.3)
    assert np.array_equal(full["noiseLevel"], np.eye(n))               # DiracKernel.scala:49-53: value / |s| on |x-y| == 0
    cross = dg.SVMKernel.buildCrossKernelGradMatrix(X, Xs, hyp, k)
    wantc = ok.grads(X, Xs)
    assert entry_rel(cross["kernel-matrix"], ok.value(X, Xs)) <= TOL_K
    for h in hyp:
        scale = float(np.abs(wantc[h]).max()) or 1.0
        assert cross[h].shape == (n, m) and float(np.abs(cross[h] - wantc[h]).max()) <= TOL_K * scale, h


def test_gradient_matrices_of_a_composite_and_their_trace_identity():
    """For a sum / product expression, and the identity the trace pass rests on: sum_ij M_ij dK_ij/dtheta with the
    materialised matrices equals gradEnergy's fused value."""
    n, d = 220, 4
    X, y, _ = orc.synthetic(n, d, seed=6)
    k = dg.SEKernel(1.6, 1.1) * dg.CauchyKernel(2.0) + dg.GenericMaternKernel(2.5, 2)
    ok = orc.Sum(orc.Prod(orc.Kernel("se", {"bandwidth": 1.6, "amplitude": 1.1}), orc.Kernel("cauchy", {"sigma": 2.0})),
                 orc.Kernel("matern", {"p": 2.0, "l": 2.5}))
    noise = 0.2
    model = dg.GPRegression(k, dg.DiracKernel(noise), list(zip(X, y)))
    g = model.gradEnergy(model._current_state)
    names = [h for h in k.effective_hyper_parameters]
    mats = dg.SVMKernel.buildKernelGradMatrix(X, names + ["noiseLevel"], k, noise=noise)
    K = mats["kernel-matrix"]
    Kinv = np.linalg.inv(K)
    alpha = Kinv @ y
    M = np.outer(alpha, alpha) - Kinv
    want = ok.grads(X, X)
    for (gk, gv), h in zip(g.items(), names + ["noiseLevel"]):
        assert gk.split("/")[-1] == h.split("/")[-1]
        assert float(np.sum(M * mats[h])) == pytest.approx(gv, rel=1e-8, abs=1e-8 * max(abs(v) for v in g.values()))
    for hn, (on, ov) in zip(names, want.items()):
        assert float(np.abs(mats[hn] - ov).max()) <= TOL_K * float(np.abs(ov).max()), hn
