"""Generate the golden fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the fixtures travel to the GPU box, the reference does not).

  housing.npz   -- config C1: REF/data/housing.csv (450 x 14, MEDV last) as train, housingTest.csv (56) as
                   test, Gaussian-scaled with the training statistics as
                   TestGPHousing.scala:117-135 / DynaMLPipe.gaussianScalingTrainTest do; SE(l,a)+Dirac(s)
                   energy, gradient, posterior mean / covariance / error bars from the oracle.
  synthetic.npz -- seeded synthetic problems for every kernel: energy, gradient (both
                   grad modes for ARD), posterior mean / covariance.
  kats.json     -- the reference's own known-answer values for this path, with file:line.
  extended.npz  -- the rows of SURVEY 8(f): kernel algebra, tensor kernels, basis-function GP, extended skew GP,
                   Kronecker multi-output GP and the Student-t process on one seeded problem (oracle values).
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import gp_oracle as orc  # noqa: E402

REF = Path("/root/reference")


def housing():
    tr = np.loadtxt(REF / "data" / "housing.csv")  # whitespace separated despite the name
    te = np.loadtxt(REF / "data" / "housingTest.csv")
    Xtr, ytr, Xte, yte = tr[:, :13], tr[:, 13], te[:, :13], te[:, 13]
    # gaussianScalingTrainTest: features and targets standardised with the training mean / sample std
    mx, sx = Xtr.mean(0), Xtr.std(0, ddof=1)
    my, sy = ytr.mean(), ytr.std(ddof=1)
    Xtr, Xte = (Xtr - mx) / sx, (Xte - mx) / sx
    ytr, yte = (ytr - my) / sy, (yte - my) / sy
    hyp = {"bandwidth": 2.5, "amplitude": 1.5}
    noise = 0.5
    k = orc.Kernel("se", hyp)
    e = orc.energy(k, noise, Xtr, ytr)
    g = orc.grad_energy(k, noise, Xtr, ytr)
    mu, cov = orc.predictive(k, noise, Xtr, ytr, Xte)
    lo, hi = orc.error_bars(mu, cov, 2.0)
    K = orc.build_kernel_matrix(k, Xtr, orc.dirac(noise))
    np.savez_compressed(HERE / "housing.npz", Xtr=Xtr, ytr=ytr, Xte=Xte, yte=yte, bandwidth=hyp["bandwidth"],
                        amplitude=hyp["amplitude"], noise=noise, energy=e,
                        grad_names=np.array(list(g.keys())), grad=np.array(list(g.values())), mu=mu, cov=cov, lo=lo,
                        hi=hi, sigma=2.0, K_first_col=K[:, 0], K_diag=np.diag(K))


def synthetic():
    out = {}
    n, d, m = 384, 6, 40
    X, y, Xs = orc.synthetic(n, d, seed=7, m=m)
    out["X"], out["y"], out["Xs"] = X, y, Xs
    band = np.linspace(1.5, 3.0, d)
    kernels = {
        "se": orc.Kernel("se", {"bandwidth": 1.7, "amplitude": 1.2}),
        "rbf": orc.Kernel("rbf", {"bandwidth": 2.1}),
        "ard_ref": orc.Kernel("ard", {"MahalanobisAmplitude": 1.1,
                                      **{f"MahalanobisBandwidth_{i}": b for i, b in enumerate(band)}}, "reference"),
        "ard_exact": orc.Kernel("ard", {"MahalanobisAmplitude": 1.1,
                                        **{f"MahalanobisBandwidth_{i}": b for i, b in enumerate(band)}}, "exact"),
        "matern1": orc.Kernel("matern", {"p": 1.0, "l": 2.2}),
        "matern2": orc.Kernel("matern", {"p": 2.0, "l": 2.2}),
        "matern3": orc.Kernel("matern", {"p": 3.0, "l": 2.2}),
        "periodic": orc.Kernel("periodic", {"sigma": 1.3, "lengthscale": 1.4, "frequency": 0.07}),
        # SURVEY 8(f) rank 1
        "cauchy": orc.Kernel("cauchy", {"sigma": 2.3}),
        "laplacian": orc.Kernel("laplacian", {"beta": 3.1}),
        "ratquad": orc.Kernel("ratquad", {"mu": 1.5, "l": 1.9}),
        "tstudent": orc.Kernel("tstudent", {"d": 1.6}),
    }
    noise = 0.25
    out["noise"] = noise
    out["band"] = band
    X_all, Xs_all = X, Xs
    for name, k in kernels.items():
        # the periodic kernel takes |x-y|_1 (Q5), which is only guaranteed PSD for 1-D inputs
        X, Xs = (X_all[:, :1], Xs_all[:, :1]) if name == "periodic" else (X_all, Xs_all)
        out[f"{name}_energy"] = orc.energy(k, noise, X, y)
        g = orc.grad_energy(k, noise, X, y)
        out[f"{name}_grad_names"] = np.array(list(g.keys()))
        out[f"{name}_grad"] = np.array(list(g.values()))
        mu, cov = orc.predictive(k, noise, X, y, Xs)
        out[f"{name}_mu"], out[f"{name}_cov"] = mu, cov
        out[f"{name}_Kcol"] = orc.build_kernel_matrix(k, X)[:, 3]
    np.savez_compressed(HERE / "synthetic.npz", **out)


def kats():
    import math
    rho = math.exp(-1.0 / 2)
    k = {
        "gp_energy_identity": {"value": -math.log(0.15915494309189535), "eps": 1e-5,
                               "ref": "dynaml-core/src/test/scala/io/github/tailhq/dynaml/models/gp/GPSpec.scala:13-38"},
        "gp_posterior_mean": {"value": rho, "ref": "GPSpec.scala:40-76"},
        "gp_posterior_var": {"value": 1.0 - rho * rho, "ref": "GPSpec.scala:68"},
        "gp_bar_lo": {"value": rho - math.sqrt(1.0 - rho * rho), "ref": "GPSpec.scala:70-72"},
        "gp_bar_hi": {"value": rho + math.sqrt(1.0 - rho * rho), "ref": "GPSpec.scala:70-72"},
        "se_value_1_0": {"value": 0.6065306597126334, "eps": 1e-5,
                         "ref": "dynaml-core/src/test/scala/io/github/tailhq/dynaml/kernels/KernelSpec.scala:34-58"},
        "builder_eval2_offdiag": {"value": 0.5, "ref": "KernelSpec.scala:181-231"},
        "laplace_value_1_0": {"value": 0.36787944117144233, "eps": 1e-5, "ref": "KernelSpec.scala:34-58"},
        "cauchy_value_1_1p5": {"value": 0.8, "eps": 1e-5, "ref": "KernelSpec.scala:34-58"},
        "bcholesky_diag_tol": {"value": 1e-6,
                               "ref": "dynaml-core/src/test/scala/io/github/tailhq/dynaml/algebra/PartitionedLinearAlgebraSpec.scala:14-46"},
    }
    (HERE / "kats.json").write_text(json.dumps(k, indent=1))


def extended_basis(x):
    return np.array([1.0, x[0], x[1] * x[2]])


EXT_B = np.array([0.3, 0.5, -0.2])
EXT_COVB = np.diag([0.5, 0.2, 0.3]) + 0.02


def extended_kernels():
    se = lambda: orc.Kernel("se", {"bandwidth": 1.4, "amplitude": 1.1})  # noqa: E731
    cau = lambda: orc.Kernel("cauchy", {"sigma": 2.0})  # noqa: E731
    mat = lambda: orc.Kernel("matern", {"p": 2.0, "l": 1.7})  # noqa: E731
    lap = lambda: orc.Kernel("laplacian", {"beta": 2.0})  # noqa: E731
    return {
        "algebra": orc.Sum(orc.Prod(se(), cau()), orc.Scaled(mat(), 0.6)),          # se * cauchy + matern * 0.6
        "tensor": orc.Prod(orc.Sliced(se(), 0, 2), orc.Sliced(lap(), 2)),           # se (x) laplacian, split 2
    }


def extended():
    X, y, Xs = orc.synthetic(220, 4, seed=77, m=12)
    noise = 0.3
    out = {"X": X, "y": y, "Xs": Xs, "noise": noise}
    for name, k in extended_kernels().items():
        out[f"{name}_energy"] = orc.energy(k, noise, X, y)
        g = orc.grad_energy(k, noise, X, y)
        out[f"{name}_grad"] = np.array(list(g.values()))
        mu, cov = orc.predictive(k, noise, X, y, Xs)
        out[f"{name}_mu"], out[f"{name}_cov"] = mu, cov
    se = orc.Kernel("se", {"bandwidth": 1.6, "amplitude": 1.1})
    mean = np.array([extended_basis(x) @ EXT_B for x in X])
    out["basis_energy"] = orc.energy(orc.WithFeatureMap(se, extended_basis, EXT_COVB), noise, X, y, mean)
    out["esgp_energy"] = orc.esgp_energy(se, noise, X, y, 0.6, 0.3)
    mu, cov, skew, cutoff = orc.esgp_predictive(se, noise, X, y, Xs, 0.6, 0.3)
    out["esgp_mu"], out["esgp_skew"], out["esgp_cutoff"] = mu, skew, cutoff
    rng = np.random.default_rng(78)
    Y = np.column_stack([y, np.cos(X[:, 0]) + 0.1 * rng.standard_normal(len(y)), 0.5 * y + 0.2 * rng.standard_normal(len(y))])
    out["Y"] = Y
    out["mogp_energy"] = orc.kronecker_mogp_energy(orc.Kernel("matern", {"p": 2.0, "l": 1.7}), 0.4,
                                                   orc.Kernel("cauchy", {"sigma": 2.0}), X, Y)
    out["stp_energy"] = orc.stp_energy(4.5, se, noise, X, y)
    np.savez_compressed(HERE / "extended.npz", **out)


if __name__ == "__main__":
    housing()
    synthetic()
    kats()
    extended()
    print("golden fixtures written to", HERE)
