"""Mid-N chain timings (C2 / C5 regime): phases of energy and energy + gradient at N = 8192 / 16384 / 32768, the
chained triangular solves against the per-block ones (bit-identical results required), and evaluations/s.
usage: midn.py [sizes, comma separated]"""
import sys, time, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
sizes = [int(s) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["8192", "16384", "32768"])]


def synthetic(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)); w = rng.standard_normal(d) / np.sqrt(d)
    return X, np.sin(X @ w) + 0.1 * rng.standard_normal(n)


for n in sizes:
    d = 8
    X, y = synthetic(n, d)
    data = ctx.data_create(X, y)
    spec = dg.MahalanobisKernel([float(np.sqrt(d))] * d, 1.0)._spec(noise=0.1)
    rec = {"n": n}
    ref = {}
    for chained in (0, 1):
        ctx.set_option("trsv_chained", chained)
        for _ in range(2):
            e = ctx.energy(data, spec)
        t0 = time.perf_counter()
        reps = 10 if n <= 16384 else 3
        for _ in range(reps):
            e = ctx.energy(data, spec)
        dt_e = (time.perf_counter() - t0) / reps
        te = ctx.timings()
        for _ in range(2):
            eg, g = ctx.energy_grad(data, spec)
        t0 = time.perf_counter()
        for _ in range(reps):
            eg, g = ctx.energy_grad(data, spec)
        dt_g = (time.perf_counter() - t0) / reps
        tg = ctx.timings()
        ref[chained] = (e, eg, np.asarray(g))
        rec[f"chained={chained}"] = {
            "energy_ms": round(dt_e * 1e3, 3), "energy_phases": {k: round(v, 3) for k, v in te.items() if v},
            "energy_grad_ms": round(dt_g * 1e3, 3), "evals_per_s": round(1 / dt_g, 2),
            "grad_phases": {k: round(v, 3) for k, v in tg.items() if v},
            "potrf_tflops": round(n ** 3 / 3 / (tg["potrf"] * 1e-3) / 1e12, 2)}
    rec["bit_identical"] = bool(ref[0][0] == ref[1][0] and ref[0][1] == ref[1][1] and np.array_equal(ref[0][2], ref[1][2]))
    rec["energy"] = ref[1][0]
    print(json.dumps(rec), flush=True)
    ctx.data_destroy(data)
    ctx.trim()
