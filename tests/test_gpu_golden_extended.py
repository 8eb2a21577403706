"""Device results against the committed fixture tests/golden/extended.npz (the SURVEY 8f rows: kernel algebra, tensor
kernels, basis-function GP, extended skew GP, Kronecker multi-output GP, Student-t process)."""
from pathlib import Path

import numpy as np
import pytest

import dynaml_b200 as dg

pytestmark = pytest.mark.gpu
Z = np.load(Path(__file__).parent / "golden" / "extended.npz")
TOL = 1e-9
X, y, Xs, NOISE = Z["X"], Z["y"], Z["Xs"], float(Z["noise"])


def _kernels():
    se = lambda: dg.SEKernel(1.4, 1.1)  # noqa: E731
    return {"algebra": se() * dg.CauchyKernel(2.0) + dg.GenericMaternKernel(1.7, 2) * 0.6,
            "tensor": dg.KroneckerProductKernel(se(), dg.LaplacianKernel(2.0), 2)}


@pytest.mark.parametrize("name", ["algebra", "tensor"])
def test_composite_kernels_against_fixture(name):
    model = dg.GPRegression(_kernels()[name], dg.DiracKernel(NOISE), list(zip(X, y)))
    h = model._current_state
    assert model.energy(h) == pytest.approx(float(Z[f"{name}_energy"]), rel=TOL)
    g = np.array(list(model.gradEnergy(h).values()))
    ref = Z[f"{name}_grad"]
    assert np.abs(g - ref).max() <= TOL * np.abs(ref).max()
    post = model.predictiveDistribution(Xs)
    assert np.abs(post.mu - Z[f"{name}_mu"]).max() <= TOL * np.abs(Z[f"{name}_mu"]).max()
    assert np.abs(post.covariance - Z[f"{name}_cov"]).max() <= TOL * np.abs(Z[f"{name}_cov"]).max()


def test_sibling_models_against_fixture():
    basis = lambda x: np.array([1.0, x[0], x[1] * x[2]])  # noqa: E731
    b, covB = np.array([0.3, 0.5, -0.2]), np.diag([0.5, 0.2, 0.3]) + 0.02
    bm = dg.GPBasisFuncRegression(dg.SEKernel(1.6, 1.1), dg.DiracKernel(NOISE), list(zip(X, y)), basis, (b, covB))
    assert bm.energy(bm._current_state) == pytest.approx(float(Z["basis_energy"]), rel=TOL)
    es = dg.ESGPModel(dg.SEKernel(1.6, 1.1), dg.DiracKernel(NOISE), list(zip(X, y)), 0.6, 0.3)
    assert es.energy(es._current_state) == pytest.approx(float(Z["esgp_energy"]), rel=TOL)
    post = es.predictiveDistribution(Xs)
    assert np.abs(post.mu - Z["esgp_mu"]).max() <= TOL * np.abs(Z["esgp_mu"]).max()
    assert np.abs(post.alpha - Z["esgp_skew"]).max() <= TOL * np.abs(Z["esgp_skew"]).max()
    assert post.tau == pytest.approx(float(Z["esgp_cutoff"]), rel=TOL)
    mo = dg.KroneckerMOGPModel(dg.GenericMaternKernel(1.7, 2), dg.DiracKernel(0.4), dg.CoRegCauchyKernel(2.0),
                               list(zip(X, Z["Y"])), 3)
    assert mo.energy(mo._current_state) == pytest.approx(float(Z["mogp_energy"]), rel=TOL)
    stp = dg.STPRegression(2.5, dg.SEKernel(1.6, 1.1), dg.DiracKernel(NOISE), list(zip(X, y)))
    h = dict(stp._current_state)
    h["degrees_of_freedom"] = 2.5   # energy evaluates the density at h["degrees_of_freedom"] + 2 (reference quirk)
    assert stp.energy(h) == pytest.approx(float(Z["stp_energy"]), rel=TOL)
