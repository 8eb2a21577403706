"""CPU baseline for bench.py -- TEST/MEASUREMENT INFRASTRUCTURE ONLY (see oracle/gp_oracle.py).

The reference (Scala on Breeze/netlib) cannot run in this image, so the timed CPU arm is this
port ("kind": "port"): the same math as gp_oracle.energy / grad_energy, formulated the way a
LAPACK user would (dpotrf + dpotri and elementwise traces, O(N^3) once -- not the reference's
O(p N^3) multi-solve per hyper-parameter, which would only make the CPU look worse), on all host
threads OpenBLAS is given.  Results are checked against gp_oracle in tests/test_cpu_baseline.py.

Timing protocol (`time_parts`): the O(N^2) part (assembly by direct differences, the gradient
traces) is timed on one full evaluation at `n2`; the O(N^3) part (dpotrf, the two triangular
solves, dpotri) is timed at `n3` >= 16384 on the kernel matrix of the same workload.  That matrix is
built for the timing only, through the Gram form X X' (LAPACK's run time does not depend on the
values, and direct differences at n3 would take minutes of NumPy time that is not being measured).
The two parts are scaled to the workload's N by their own orders; the result is labelled
`extrapolated` wherever it is reported.
"""
from __future__ import annotations

import math
import os
import time
from typing import Dict, Tuple

import numpy as np
import scipy.linalg as sla
from scipy.linalg import lapack

from . import gp_oracle as orc


def host_cores() -> int:
    """Cores this process may run on (cgroup / affinity aware)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def use_all_cores() -> int:
    """Give the BLAS/LAPACK thread pools every usable core, whatever OMP_NUM_THREADS said at import
    (torchrun exports OMP_NUM_THREADS=1 to its workers).  Returns the thread count in effect."""
    n = host_cores()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n, user_api="blas")
    except Exception:
        pass
    return blas_threads()


def energy_grad_fast(cov: orc.Kernel, noise_level: float, X: np.ndarray, y: np.ndarray, timing: dict | None = None
                     ) -> Tuple[float, Dict[str, float]]:
    """E and g_theta = sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij/dtheta  (same values as
    gp_oracle.energy / grad_energy up to rounding).  `timing`, if given, receives the seconds spent
    in the O(N^3) LAPACK part ("n3") and in the O(N^2) assembly / trace part ("n2")."""
    t0 = time.perf_counter()
    nz = orc.dirac(noise_level)
    K = orc.build_kernel_matrix(cov, X, nz)
    t1 = time.perf_counter()
    c, info = lapack.dpotrf(K, lower=1, overwrite_a=0)
    if info != 0 or not np.isfinite(np.diag(c)).all():
        # Breeze: NotConvergedException from dpotrf, or MatrixNotSymmetricException for a NaN matrix (NaN != NaN);
        # both are caught by the model (AbstractGPRegressionModel.scala:243-251, 531-533)
        names = cov.effective() + ["noiseLevel"]
        return float("inf"), {n: float("nan") for n in names}
    alpha = sla.cho_solve((c, True), y)
    e = 0.5 * (float(y @ alpha) + float(np.log(np.diag(c)).sum()) + len(y) * math.log(2 * math.pi))
    kinv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
    t2 = time.perf_counter()
    kinv = np.tril(kinv) + np.tril(kinv, -1).T
    M = np.outer(alpha, alpha) - kinv
    g = {n: float(np.sum(M * dK)) for n, dK in cov.grads(X, X).items() if n in cov.effective()}
    g["noiseLevel"] = float(np.sum(M * nz.grads(X, X)["noiseLevel"]))
    t3 = time.perf_counter()
    if timing is not None:
        timing["n3"] = t2 - t1
        timing["n2"] = (t1 - t0) + (t3 - t2)
    return e, g


def make_kernel(kind: str, d: int) -> orc.Kernel:
    if kind == "matern":
        return orc.Kernel("matern", {"p": 2.0, "l": math.sqrt(d)})
    if kind == "ard":
        return orc.Kernel("ard", {"MahalanobisAmplitude": 1.0,
                                  **{f"MahalanobisBandwidth_{i}": math.sqrt(d) for i in range(d)}})
    return orc.Kernel("se", {"bandwidth": math.sqrt(d), "amplitude": 1.0})


def timing_matrix(n: int, d: int, kind: str, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(K, y) of the workload at size n for timing the LAPACK part: distances through the Gram form
    (one dgemm), the workload's kernel and noise on top.  Not used for any parity statement."""
    X, y, _ = orc.synthetic(n, d, seed)
    ell = math.sqrt(d)
    sq = (X * X).sum(axis=1)
    K = X @ X.T
    K *= -2.0
    K += sq[:, None]
    K += sq[None, :]
    np.maximum(K, 0.0, out=K)
    if kind == "matern":
        np.sqrt(K, out=K)
        K *= math.sqrt(5.0) / ell
        poly = 1.0 + K + K * K / 3.0
        np.exp(-K, out=K)
        K *= poly
        del poly
    else:
        K *= -0.5 / (ell * ell)
        np.exp(K, out=K)
    K[np.diag_indices(n)] += 0.1
    return K, y


def time_lapack_part(n: int, d: int, kind: str, seed: int = 0) -> float:
    """Seconds of dpotrf + the two triangular solves + dpotri at size n (the O(N^3) part of one
    energy + gradient evaluation)."""
    K, y = timing_matrix(n, d, kind, seed)
    t0 = time.perf_counter()
    c, info = lapack.dpotrf(K, lower=1, overwrite_a=1)
    if info != 0:
        raise RuntimeError("timing matrix is not positive definite")
    sla.cho_solve((c, True), y)
    lapack.dpotri(c, lower=1, overwrite_c=1)
    return time.perf_counter() - t0


def time_energy_grad(n: int, d: int, kind: str = "matern", seed: int = 0) -> Tuple[float, float, int]:
    """(seconds in the O(N^2) part, seconds in the O(N^3) part, BLAS threads) of one energy+grad
    evaluation at size n."""
    X, y, _ = orc.synthetic(n, d, seed)
    tm: dict = {}
    energy_grad_fast(make_kernel(kind, d), 0.1, X, y, tm)
    return tm["n2"], tm["n3"], blas_threads()


def time_parts(n2: int, n3: int, d: int, kind: str = "matern", seed: int = 0) -> dict:
    """One sample of the CPU arm: the O(N^2) part measured at n2 (full evaluation), the O(N^3) part at n3."""
    t2, t3_small, threads = time_energy_grad(n2, d, kind, seed)
    if n3 > n2:
        t3 = time_lapack_part(n3, d, kind, seed)
    else:
        n3, t3 = n2, t3_small
    return {"n2": n2, "t2": t2, "n3": n3, "t3": t3, "threads": threads}


def extrapolate(t_n2: float, t_n3: float, n_sample: int, n_full: int, n3_sample: int | None = None) -> float:
    """Seconds per evaluation at n_full: each part scaled from its own sample size by its own order."""
    r2 = n_full / n_sample
    r3 = n_full / (n3_sample or n_sample)
    return t_n2 * r2 * r2 + t_n3 * r3 * r3 * r3


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        ts = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if ts:
            return int(max(ts))
    except Exception:
        pass
    return int(os.environ.get("OPENBLAS_NUM_THREADS", os.cpu_count() or 1))
