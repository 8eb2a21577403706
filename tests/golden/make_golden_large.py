"""Oracle values at the BASELINE configuration sizes -> tests/golden/large.json.

The oracle needs minutes of CPU time and tens of GB of host memory at these sizes, so it is run once, here in the
build container, and the values are committed; the GPU tests (tests/test_gpu_large.py) and bench.py's parity legs
compare against them on the same seeded inputs (oracle.gp_oracle.synthetic).  The arithmetic is the oracle's:
kernel entries by direct differences (orc.Kernel.value / .grads, evaluated in row blocks to bound memory), LAPACK
dpotrf / dtrtrs / dpotri through SciPy, the energy and gradient formulas of gp_oracle.energy / grad_energy in the
dpotri formulation of oracle/cpu_baseline.py (tests/test_cpu_baseline.py pins that formulation to gp_oracle).

    python tests/golden/make_golden_large.py [case ...]       # default: every case
"""
import json
import math
import sys
import time
from pathlib import Path

import numpy as np
import scipy.linalg as sla
from scipy.linalg import lapack

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import gp_oracle as orc  # noqa: E402

OUT = HERE / "large.json"
BLOCK = 1024


def kernel_matrix(cov, X, noise):
    """orc.build_kernel_matrix(cov, X, dirac(noise)) in row blocks (same per-entry arithmetic)."""
    n = X.shape[0]
    K = np.empty((n, n))
    for i0 in range(0, n, BLOCK):
        K[i0:i0 + BLOCK, :] = cov.value(X[i0:i0 + BLOCK], X)
    # DiracKernel: |noise| wherever |x - y|_2 == 0 -- for pairwise distinct points, the diagonal
    assert len(np.unique(X, axis=0)) == n
    K[np.diag_indices(n)] += abs(noise)
    return K


BIG = 4096   # block edge of the out-of-LAPACK path below


def _chol_solve_inverse_blocked(K, y, want_inv):
    """N >= 32768: both OpenBLAS builds in this image (SciPy's LP64 and NumPy's ILP64) crash inside their threaded
    dpotrf / dpotri at this size, so the factorisation, the triangular solves and the inverse are blocked by hand:
    LAPACK (dpotrf, dtrtrs) on BIG x BIG diagonal blocks only, dgemm for everything else.
    Returns (L, z = L^-1 y, alpha = L^-T z, lower triangle of K^-1 or None)."""
    n = K.shape[0]
    nb = range(0, n, BIG)
    # right-looking blocked Cholesky in place (lower): dpotrf on BIG x BIG diagonal blocks, dtrsm / dgemm around them
    L = K
    for j in nb:
        L[j:j + BIG, j:j + BIG] = sla.cholesky(L[j:j + BIG, j:j + BIG], lower=True)
        if j + BIG < n:
            L[j + BIG:, j:j + BIG] = sla.solve_triangular(L[j:j + BIG, j:j + BIG], L[j + BIG:, j:j + BIG].T, lower=True).T
            for i in range(j + BIG, n, BIG):
                L[i:i + BIG, j + BIG:i + BIG] -= L[i:i + BIG, j:j + BIG] @ L[j + BIG:i + BIG, j:j + BIG].T
    for j in nb:   # the strictly upper blocks still hold K: clear them
        L[j:j + BIG, j + BIG:] = 0.0
        L[j:j + BIG, j:j + BIG] = np.tril(L[j:j + BIG, j:j + BIG])
    z = y.astype(np.float64).copy()
    for j in nb:
        z[j:j + BIG] = sla.solve_triangular(L[j:j + BIG, j:j + BIG], z[j:j + BIG], lower=True)
        z[j + BIG:] -= L[j + BIG:, j:j + BIG] @ z[j:j + BIG]
    alpha = z.copy()
    for j in reversed(nb):
        alpha[j:j + BIG] = sla.solve_triangular(L[j:j + BIG, j:j + BIG], alpha[j:j + BIG], lower=True, trans="T")
        alpha[:j] -= L[j:j + BIG, :j].T @ alpha[j:j + BIG]
    if not want_inv:
        return L, z, alpha, None
    W = np.zeros_like(L)                       # W = L^-1, block column by block column
    eye = np.eye(BIG)
    for j in nb:
        W[j:j + BIG, j:j + BIG] = sla.solve_triangular(L[j:j + BIG, j:j + BIG], eye[:min(BIG, n - j), :min(BIG, n - j)], lower=True)
        for i in range(j + BIG, n, BIG):
            acc = L[i:i + BIG, j:i] @ W[j:i, j:j + BIG]
            W[i:i + BIG, j:j + BIG] = -sla.solve_triangular(L[i:i + BIG, i:i + BIG], acc, lower=True)
    kinv = np.zeros_like(L)                    # lower blocks of W' W
    for j in nb:
        for i in range(j, n, BIG):
            kinv[i:i + BIG, j:j + BIG] = W[i:, i:i + BIG].T @ W[i:, j:j + BIG]
    return L, z, alpha, kinv


def energy_and_grad(cov, noise, X, y, want_grad=True, blocked=None):
    n = X.shape[0]
    t0 = time.perf_counter()
    K = kernel_matrix(cov, X, noise)
    if blocked is None:
        blocked = n >= 32768
    if blocked:
        c, z, alpha_full, kinv_b = _chol_solve_inverse_blocked(K, y, want_grad)
        del K
        alpha = z
    else:
        c, info = lapack.dpotrf(K, lower=1, overwrite_a=1)
        assert info == 0
        del K
        alpha = sla.solve_triangular(c, y, lower=True)
    quad = float(alpha @ alpha)
    logdet_half = float(np.log(np.diag(c)).sum())
    e = 0.5 * (quad + logdet_half + n * math.log(2 * math.pi))
    out = {"energy": e, "quad": quad, "logdet_half": logdet_half}
    if want_grad:
        if blocked:
            alpha, kinv = alpha_full, kinv_b
        else:
            alpha = sla.solve_triangular(c, alpha, lower=True, trans="T")
            kinv, info = lapack.dpotri(c, lower=1, overwrite_c=1)  # lower triangle of K^-1
            assert info == 0
        names = cov.effective()
        g = {k: 0.0 for k in names}
        for i0 in range(0, n, BLOCK):
            i1 = min(n, i0 + BLOCK)
            blk = kinv[i0:i1, :i1].copy()
            # symmetric M = alpha alpha' - K^-1 on rows i0:i1, columns 0:i1 (lower part), strict lower counted twice
            Mb = np.outer(alpha[i0:i1], alpha[:i1]) - blk
            Mb[:, :i0] *= 2.0
            D = Mb[:, i0:i1]
            D[:] = 2.0 * np.tril(D, -1) + np.diag(np.diag(D))
            dK = cov.grads(X[i0:i1], X[:i1])
            for k in names:
                g[k] += float(np.sum(Mb * dK[k]))
        # noise: d|s|/ds on the diagonal (DiracKernel.scala:49-53: value / |s| = 1 wherever |x-y| == 0)
        g["noiseLevel"] = float(np.sum(alpha * alpha) - np.trace(kinv))
        out["grad"] = g
    out["seconds"] = time.perf_counter() - t0
    return out


def ard_kernel(d, scale=1.0):
    return orc.Kernel("ard", {"MahalanobisAmplitude": 1.0,
                              **{f"MahalanobisBandwidth_{i}": math.sqrt(d) * scale * (1.0 + 0.05 * i) for i in range(d)}})


CASES = {}


def case(f):
    CASES[f.__name__] = f
    return f


@case
def c4_check():
    """bench.py's distributed-Cholesky parity leg: n = 8192, d = 8, SE(sqrt 8, 1), noise 0.1, seed 0."""
    X, y, _ = orc.synthetic(8192, 8, seed=0)
    r = energy_and_grad(orc.Kernel("se", {"bandwidth": math.sqrt(8), "amplitude": 1.0}), 0.1, X, y, want_grad=False)
    return {"n": 8192, "d": 8, "seed": 0, "kernel": "se", "bandwidth": math.sqrt(8), "amplitude": 1.0, "noise": 0.1, **r}


@case
def c2():
    """Config C2: N = 8192, d = 8, ARD-SE, energy + gradient (reference grad mode, D2)."""
    X, y, _ = orc.synthetic(8192, 8, seed=0)
    r = energy_and_grad(ard_kernel(8), 0.1, X, y)
    return {"n": 8192, "d": 8, "seed": 0, "kernel": "ard", "noise": 0.1,
            "bandwidths": [math.sqrt(8) * (1.0 + 0.05 * i) for i in range(8)], "amplitude": 1.0, **r}


@case
def se_stress():
    """the worst-conditioned configuration of the tolerance contract: SE, sigma = 1e-3, N = 8192."""
    X, y, _ = orc.synthetic(8192, 8, seed=0)
    r = energy_and_grad(orc.Kernel("se", {"bandwidth": math.sqrt(8), "amplitude": 1.0}), 1e-3, X, y)
    return {"n": 8192, "d": 8, "seed": 0, "kernel": "se", "bandwidth": math.sqrt(8), "amplitude": 1.0, "noise": 1e-3, **r}


@case
def c5():
    """Config C5: N = 16384, d = 8, SE; three candidates of bench.py's 16 x 16 GridSearch grid (first, middle, last)."""
    X, y, _ = orc.synthetic(16384, 8, seed=0)
    init = {"bandwidth": 2.5 * math.sqrt(8), "noiseLevel": 0.6}
    grid = orc.get_grid(init, 16, 0.08, True)
    out = []
    for idx in (0, 119, 255):
        cfg = grid[idx]
        r = energy_and_grad(orc.Kernel("se", {"bandwidth": cfg["bandwidth"], "amplitude": 1.0}), cfg["noiseLevel"], X, y,
                            want_grad=False)
        out.append({"index": idx, **cfg, **r})
    return {"n": 16384, "d": 8, "seed": 0, "kernel": "se", "amplitude": 1.0, "init": init, "gridsize": 16, "step": 0.08,
            "log_scale": True, "candidates": out}


@case
def c3():
    """Config C3: N = 32768, d = 16, Matern-5/2 (p = 2, l = 4), noise 0.1: energy + gradient."""
    X, y, _ = orc.synthetic(32768, 16, seed=0)
    r = energy_and_grad(orc.Kernel("matern", {"p": 2.0, "l": 4.0}), 0.1, X, y)
    return {"n": 32768, "d": 16, "seed": 0, "kernel": "matern", "p": 2, "l": 4.0, "noise": 0.1, **r}


@case
def c2_predict():
    """Posterior at a BASELINE size: the C2 model (N = 8192, d = 8, ARD-SE, noise 0.1) at 192 fresh test points:
    mean, covariance (diagonal, first 8 x 8 block, one far-apart pair), error bars at sigma = 2 (gp_oracle.predictive /
    error_bars: dpotrf + dtrtrs)."""
    X, y, Xs = orc.synthetic(8192, 8, seed=0, m=192)
    cov = ard_kernel(8)
    t0 = time.perf_counter()
    K = kernel_matrix(cov, X, 0.1)
    c, info = lapack.dpotrf(K, lower=1, overwrite_a=1)
    assert info == 0
    del K
    alpha = sla.solve_triangular(c, sla.solve_triangular(c, y, lower=True), lower=True, trans="T")
    Ks = cov.value(X, Xs)
    v = sla.solve_triangular(c, Ks, lower=True)
    post = cov.value(Xs, Xs) - v.T @ v
    mu = Ks.T @ alpha
    lo, hi = orc.error_bars(mu, post, 2.0)
    return {"n": 8192, "d": 8, "seed": 0, "m": 192, "kernel": "ard", "noise": 0.1, "sigma": 2.0,
            "bandwidths": [math.sqrt(8) * (1.0 + 0.05 * i) for i in range(8)], "amplitude": 1.0,
            "mean": mu.tolist(), "var": np.diag(post).tolist(), "cov_block": post[:8, :8].tolist(),
            "cov_far": float(post[191, 3]), "lo": lo.tolist(), "hi": hi.tolist(), "seconds": time.perf_counter() - t0}


def selftest():
    """The blocked formulation against gp_oracle.energy / grad_energy at a size the oracle does directly."""
    global BLOCK
    BLOCK = 100
    X, y, _ = orc.synthetic(333, 5, seed=2)
    for cov in (ard_kernel(5), orc.Kernel("matern", {"p": 2.0, "l": 2.0}), orc.Kernel("se", {"bandwidth": 2.0, "amplitude": 1.3})):
        r = energy_and_grad(cov, 0.05, X, y)
        eo, go = orc.energy(cov, 0.05, X, y), orc.grad_energy(cov, 0.05, X, y)
        assert abs(r["energy"] - eo) <= 1e-11 * abs(eo), (r["energy"], eo)
        scale = max(abs(v) for v in go.values())
        for k, v in go.items():
            assert abs(r["grad"][k] - v) <= 1e-9 * scale, (k, r["grad"][k], v)
    # the hand-blocked path used at N >= 32768 against the LAPACK one
    global BIG
    BIG = 128
    cov = orc.Kernel("matern", {"p": 2.0, "l": 2.0})
    a, b = energy_and_grad(cov, 0.05, X, y, blocked=False), energy_and_grad(cov, 0.05, X, y, blocked=True)
    assert abs(a["energy"] - b["energy"]) <= 1e-13 * abs(a["energy"])
    for k in a["grad"]:
        assert abs(a["grad"][k] - b["grad"][k]) <= 1e-10 * max(abs(v) for v in a["grad"].values()), k
    BIG = 4096
    BLOCK = 1024
    print("selftest ok", flush=True)


def main():
    selftest()
    want = sys.argv[1:] or list(CASES)
    data = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name in want:
        t0 = time.perf_counter()
        data[name] = CASES[name]()
        OUT.write_text(json.dumps(data, indent=1, sort_keys=True))
        print(f"{name}: {time.perf_counter() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
