"""The 2D block-cyclic distributed Cholesky + NLL (dist.cu) on ONE GPU: with world == 1 the library runs
every rank of a Pr x Pc grid on the same device (broadcasts become copies), so the block-cyclic index
logic, the panel layout and the replicated forward substitution are parity-tested against the oracle
and against the dense single-GPU path.  Real multi-GPU runs: scripts/dist_run.py under torchrun."""
import math

import numpy as np
import pytest

import dynaml_b200 as dg
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture()
def dctx():
    c = dg.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (3, 2)])
@pytest.mark.parametrize("n,tile", [(1500, 256), (2048, 512), (300, 256)])
def test_emulated_grid_matches_oracle_and_dense_path(dctx, grid, n, tile):
    X, y, _ = orc.synthetic(n, 5, seed=n + grid[0] * 10 + grid[1])
    k, ok = dg.SEKernel(2.0, 1.2), orc.Kernel("se", {"bandwidth": 2.0, "amplitude": 1.2})
    data = dctx.data_create(X, y)
    try:
        dctx.dist_init(None, 0, 1, grid[0], grid[1])
        e, ms = dctx.dist_energy(data, k._spec(noise=0.3), tile)
        dense = dctx.energy(data, k._spec(noise=0.3))
        assert e == pytest.approx(orc.energy(ok, 0.3, X, y), rel=TOL)
        assert e == pytest.approx(dense, rel=1e-12)
        assert ms > 0
    finally:
        dctx.dist_finalize()
        dctx.data_destroy(data)


def test_distributed_result_is_independent_of_the_grid(dctx):
    X, y, _ = orc.synthetic(2304, 8, seed=3)
    spec = dg.GenericMaternKernel(2.5, 2)._spec(noise=0.2)
    data = dctx.data_create(X, y)
    try:
        vals = []
        for grid in [(1, 1), (2, 2), (2, 4), (4, 2)]:
            dctx.dist_init(None, 0, 1, *grid)
            vals.append(dctx.dist_energy(data, spec, 256)[0])
            dctx.dist_finalize()
        assert max(vals) - min(vals) <= 1e-11 * abs(vals[0])
    finally:
        dctx.data_destroy(data)


def test_distributed_non_pd_is_inf(dctx):
    X, y, _ = orc.synthetic(600, 2, seed=9)
    data = dctx.data_create(X, y)
    try:
        dctx.dist_init(None, 0, 1, 2, 2)
        e, _ = dctx.dist_energy(data, dg.RBFKernel(float("nan"))._spec(noise=0.1), 256)
        assert e == float("inf")
    finally:
        dctx.dist_finalize()
        dctx.data_destroy(data)


def test_single_rank_lookahead_path_at_size(dctx):
    """world == 1, 1x1 grid: the real (two-stream, look-ahead) code path; compare with the dense path."""
    n = 8192
    X, y, _ = orc.synthetic(n, 8, seed=1)
    spec = dg.SEKernel(math.sqrt(8), 1.0)._spec(noise=0.1)
    data = dctx.data_create(X, y)
    try:
        dctx.dist_init(None, 0, 1, 1, 1)
        e, ms = dctx.dist_energy(data, spec, 512)
        assert e == pytest.approx(dctx.energy(data, spec), rel=1e-11)
    finally:
        dctx.dist_finalize()
        dctx.data_destroy(data)
