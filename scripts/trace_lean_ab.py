"""Gradient trace pass: exact-difference kernel (grad_lean = 0) against the lean Gram pass (grad_lean = 1).  Prints the trace-phase time from the library's CUDA events and the largest relative difference of
the gradient against the exact-difference pass.
usage: trace_lean_ab.py [quick]"""
import sys, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
quick = len(sys.argv) > 1


def synthetic(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)); w = rng.standard_normal(d) / np.sqrt(d)
    return X, np.sin(X @ w) + 0.1 * rng.standard_normal(n)


cases = [("matern52_d16", 16, lambda d: dg.GenericMaternKernel(float(np.sqrt(d)), 2), 0.1, [5000, 16384] + ([] if quick else [32768])),
         ("matern32_d4", 4, lambda d: dg.GenericMaternKernel(2.0, 1), 0.1, [4096]),
         ("matern72_d12", 12, lambda d: dg.GenericMaternKernel(3.0, 3), 0.1, [4096]),
         ("se_d8", 8, lambda d: dg.SEKernel(float(np.sqrt(d)), 1.3), 0.1, [1000, 8192, 16384]),
         ("se_d8_stress", 8, lambda d: dg.SEKernel(float(np.sqrt(d)), 1.0), 1e-3, [8192]),
         ("ard_d8", 8, lambda d: dg.MahalanobisKernel([float(np.sqrt(d)) * (1 + 0.1 * i) for i in range(d)], 1.2), 0.1, [3000, 8192, 16384]),
         ("se_d32", 32, lambda d: dg.SEKernel(float(np.sqrt(d)), 1.0), 0.1, [4096])]
for name, d, mk, noise, sizes in cases:
    for n in sizes:
        X, y = synthetic(n, d, seed=n)
        data = ctx.data_create(X, y)
        spec = mk(d)._spec(noise=noise)
        rec = {"case": name, "n": n}
        ref = None
        for mode in (0, 1):
            ctx.set_option("grad_lean", mode)
            for _ in range(2):
                e, g = ctx.energy_grad(data, spec)
            t = [0.0] * 3
            for r in range(3):
                e, g = ctx.energy_grad(data, spec)
                t[r] = ctx.timings()["trace"]
            g = np.asarray(g)
            if mode == 0:
                ref = g
            ok = np.isfinite(ref)
            rel = float(np.max(np.abs(g[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300))) if ok.any() else 0.0
            rec[f"lean={mode}"] = {"trace_ms": round(min(t), 4), "max_rel_diff_vs_exact": rel}
        rec["grad"] = [None if v != v else float(v) for v in ref]
        print(json.dumps(rec), flush=True)
        ctx.data_destroy(data)
        ctx.trim()
ctx.set_option("grad_lean", 1)
