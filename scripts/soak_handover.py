"""Soak test of the stage hand-over in the bulk-copy GEMM: many energies (assembly + blocked Cholesky
+ solves) and energy + gradient evaluations with the C-tile prefetch off and on, every result compared bit for bit
with the cp.async kernel's.  Runs with the default hand-over (producer-side fence.proxy.async).
usage: soak_handover.py [reps at N = 8192] [reps at N = 16384]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg


def synthetic(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)); w = rng.standard_normal(d) / np.sqrt(d)
    return X, np.sin(X @ w) + 0.1 * rng.standard_normal(n)


ctx = dg.default_context()
plan = ((8192, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 300), (16384, 8, int(sys.argv[2]) if len(sys.argv) > 2 else 40))
bad = 0
for n, d, reps in plan:
    X, y = synthetic(n, d, seed=n)
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(float(np.sqrt(d)), 1.0)._spec(noise=0.05)
    ctx.set_option("gemm_variant", 2)
    e_ref = ctx.energy(data, spec)
    eg_ref, g_ref = ctx.energy_grad(data, spec)
    ctx.set_option("gemm_variant", 4)
    t0 = time.time()
    for rep in range(reps):
        ctx.set_option("gemm_no_prefetch", rep % 2)
        e = ctx.energy(data, spec)
        bad += int(e != e_ref)
        if rep % 10 == 0:
            eg, g = ctx.energy_grad(data, spec)
            bad += int(eg != eg_ref or not np.array_equal(np.asarray(g), np.asarray(g_ref)))
    ctx.set_option("gemm_no_prefetch", 0)
    print(f"N={n}: {reps} energies + {(reps + 9) // 10} energy+gradient evaluations in {time.time() - t0:.1f} s, mismatches so far {bad}", flush=True)
    ctx.data_destroy(data)
print("soak", "FAILED" if bad else "ok")
sys.exit(1 if bad else 0)
