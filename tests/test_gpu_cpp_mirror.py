"""The C++ host-side mirror (include/dynaml_b200.hpp) exercised by tests/cpp/host_mirror_test.cpp -- the
reference's GPSpec / KernelSpec cases and a seeded synthetic problem through the compiled-language binding
of the C ABI -- checked here against the oracle."""
import json
import math
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "dynaml_b200" / "lib" / "host_mirror_test"
TOL = 1e-9


@pytest.fixture(scope="module")
def result(tmp_path_factory):
    assert BIN.exists(), "build it with `python -m dynaml_b200.build` (__graft_entry__.build() does)"
    X, y, Xs = orc.synthetic(400, 4, seed=31, m=25)
    path = tmp_path_factory.mktemp("cpp") / "problem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<qqq", X.shape[0], X.shape[1], Xs.shape[0]))
        f.write(np.ascontiguousarray(X).tobytes())
        f.write(np.ascontiguousarray(y).tobytes())
        f.write(np.ascontiguousarray(Xs).tobytes())
    r = subprocess.run([str(BIN), str(path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout), (X, y, Xs)


def test_reference_known_answers_through_cpp(result):
    out, _ = result
    rho = math.exp(-0.5)
    assert out["gpspec_likelihood"] - (-math.log(0.15915494309189535)) < 1e-5      # GPSpec.scala:36
    assert out["gpspec_likelihood"] == pytest.approx(math.log(2 * math.pi), abs=1e-14)
    assert out["gpspec_pmean"] == pytest.approx(rho, rel=4e-16)                      # GPSpec.scala:68
    assert out["gpspec_psigma"] == pytest.approx(1.0 - rho * rho, rel=1e-15)
    assert out["gpspec_lo"] == pytest.approx(rho - math.sqrt(1.0 - rho * rho), abs=1e-15)
    assert out["gpspec_hi"] == pytest.approx(rho + math.sqrt(1.0 - rho * rho), abs=1e-15)
    assert abs(out["kernelspec_se"] - 0.6065306597126334) < 1e-15                    # KernelSpec.scala:53
    assert abs(out["kernelspec_laplace"] - 0.36787944117144233) < 1e-15
    assert abs(out["kernelspec_cauchy"] - 0.8) < 1e-15
    assert out["block_unblock_ok"] and out["missing_key_throws"]
    assert out["kernelspec_partitioned_ok"]                                            # KernelSpec.scala:205-229
    assert out["kernelspec_grad_sigma_01"] == pytest.approx(0.5, rel=1e-15) and out["kernelspec_grad_sigma_00"] == 0.0


def test_cpp_model_matches_oracle(result):
    out, (X, y, Xs) = result
    k = orc.Kernel("matern", {"p": 2.0, "l": 2.0})
    assert out["energy"] == pytest.approx(orc.energy(k, 0.3, X, y), rel=TOL)
    g = orc.grad_energy(k, 0.3, X, y)
    assert out["grad_keys"] == 2
    assert out["grad_l"] == pytest.approx(g["l"], rel=TOL) and out["grad_noiseLevel"] == pytest.approx(g["noiseLevel"], rel=TOL)
    mu, cov = orc.predictive(k, 0.3, X, y, Xs)
    assert np.abs(np.array(out["mu"]) - mu).max() <= TOL * np.abs(mu).max()
    got_cov = np.array(out["cov"]).reshape(len(Xs), len(Xs), order="F")
    assert np.abs(got_cov - cov).max() <= TOL * np.abs(cov).max()
    lo, hi = orc.error_bars(mu, cov, 2.0)
    assert np.abs(np.array(out["lo"]) - lo).max() <= 1e-8 * np.abs(lo).max()
    assert np.abs(np.array(out["hi"]) - hi).max() <= 1e-8 * np.abs(hi).max()


def test_cpp_tuners(result):
    out, (X, y, Xs) = result
    # grid in sorted key order (std::map): l, noiseLevel
    grid = orc.get_grid({"l": 2.6, "noiseLevel": 0.9}, 3, 0.2, False)
    ref = [orc.energy(orc.Kernel("matern", {"p": 2.0, "l": c["l"]}), c["noiseLevel"], X, y) for c in grid]
    assert np.abs(np.array(out["landscape"]) - np.array(ref)).max() <= TOL * np.abs(ref).max()
    best = grid[int(np.argmin(ref))]
    assert out["best_l"] == best["l"] and out["best_noiseLevel"] == best["noiseLevel"]
    mu, _ = orc.predictive(orc.Kernel("matern", {"p": 2.0, "l": best["l"]}), best["noiseLevel"], X, y, Xs[:1])
    assert out["persisted_predict0"] == pytest.approx(mu[0], rel=TOL)
    assert math.isfinite(out["csa_energy"]) and out["csa_energy"] <= max(ref) + 1e-9
    assert out["nonpd_energy_is_inf"] and out["nonpd_grad_is_nan"]


def test_cpp_kernel_algebra(result):
    out, (X, y, Xs) = result
    ok = orc.Sum(orc.Prod(orc.Kernel("se", {"bandwidth": 1.4, "amplitude": 1.1}), orc.Kernel("cauchy", {"sigma": 2.0})),
                 orc.Scaled(orc.Kernel("matern", {"p": 2.0, "l": 1.7}), 0.6))
    assert out["composite_energy"] == pytest.approx(orc.energy(ok, 0.3, X, y), rel=TOL)
    go = orc.grad_energy(ok, 0.3, X, y)
    ref = np.array(list(go.values()))  # bandwidth, amplitude, sigma, l, noiseLevel
    assert np.abs(np.array(out["composite_grad"]) - ref).max() <= TOL * np.abs(ref).max()
    mu, _ = orc.predictive(ok, 0.3, X, y, Xs)
    assert np.abs(np.array(out["composite_mu"]) - mu).max() <= TOL * np.abs(mu).max()


def test_cpp_basis_function_model(result):
    out, (X, y, Xs) = result
    basis = lambda x: np.array([1.0, x[0], x[1] * x[2]])  # noqa: E731
    b = np.array([0.3, 0.5, -0.2])
    covB = np.diag([0.5, 0.2, 0.3]) + 0.02
    ok = orc.WithFeatureMap(orc.Kernel("se", {"bandwidth": 1.6, "amplitude": 1.1}), basis, covB)
    mean = np.array([basis(x) @ b for x in X])
    mean_s = np.array([basis(x) @ b for x in Xs])
    assert out["basis_energy"] == pytest.approx(orc.energy(ok, 0.3, X, y, mean), rel=TOL)
    mu, _ = orc.predictive(ok, 0.3, X, y, Xs, mean, mean_s)
    assert np.abs(np.array(out["basis_mu"]) - mu).max() <= TOL * np.abs(mu).max()


def test_candidates_over_several_gpus_from_one_process():
    """tests/cpp/multi_gpu_batch_test.cpp: one context per host thread per GPU, and dgp_energy_batch_multi, both bit
    for bit equal to the single-context batch (on a one-GPU box: two contexts on that GPU)."""
    exe = ROOT / "dynaml_b200" / "lib" / "multi_gpu_batch_test"
    assert exe.exists(), "build it with `python -m dynaml_b200.build`"
    r = subprocess.run([str(exe), "1500"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout)
    assert out["mismatches"] == 0 and out["bad_arg_checks"] == 0 and out["contexts"] >= 2
    assert out["non_pd_status"] == 1          # DGP_ERR_NOT_PD for the NaN-bandwidth candidate, the others unaffected


def test_model_energy_batch_over_devices():
    """The Python mirror of the same: GPRegression.useDevices shards energyBatch / GridSearch over GPUs from this one
    process; identical energies (here every visible GPU, or device 0 alone when there is just one)."""
    import dynaml_b200 as dg
    from dynaml_b200 import _capi
    X, y, _ = orc.synthetic(900, 5, seed=4)
    se = dg.SEKernel(2.0, 1.0)
    model = dg.GPRegression(se, dg.DiracKernel(0.3), list(zip(X, y)))
    gs = dg.GridSearch(model).setGridSize(3).setStepSize(0.2)
    init = {"bandwidth": 2.5, "amplitude": 1.4, "noiseLevel": 0.6}
    single = gs.getEnergyLandscape(init)
    ndev = max(1, _capi.device_count())
    model.useDevices(list(range(ndev)) if ndev > 1 else [0, 0])   # one GPU: a second context on it
    multi = gs.getEnergyLandscape(init)
    assert [e for e, _ in multi] == [e for e, _ in single]
    model.useDevices([])
