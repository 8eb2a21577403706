"""CPU baseline for bench.py -- TEST/MEASUREMENT INFRASTRUCTURE ONLY (see oracle/gp_oracle.py).

The reference (Scala on Breeze/netlib) cannot run in this image, so the timed CPU arm is this
port ("kind": "port"): the same math as gp_oracle.energy / grad_energy, formulated the way a
LAPACK user would (dpotrf + dpotri and elementwise traces, O(N^3) once -- not the reference's
O(p N^3) multi-solve per hyper-parameter, which would only make the CPU look worse), on all host
threads OpenBLAS is given.  Results are checked against gp_oracle in tests/test_cpu_baseline.py.
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
    if info != 0:
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


def time_energy_grad(n: int, d: int, kind: str = "matern", seed: int = 0) -> Tuple[float, float, int]:
    """(seconds in the O(N^2) part, seconds in the O(N^3) part, BLAS threads) of one energy+grad
    evaluation at size n."""
    X, y, _ = orc.synthetic(n, d, seed)
    tm: dict = {}
    energy_grad_fast(make_kernel(kind, d), 0.1, X, y, tm)
    return tm["n2"], tm["n3"], blas_threads()


def extrapolate(t_n2: float, t_n3: float, n_sample: int, n_full: int) -> float:
    """Seconds per evaluation at n_full from a sample at n_sample: each part scaled by its own order."""
    r = n_full / n_sample
    return t_n2 * r * r + t_n3 * r * r * r


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        ts = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if ts:
            return int(max(ts))
    except Exception:
        pass
    return int(os.environ.get("OPENBLAS_NUM_THREADS", os.cpu_count() or 1))
