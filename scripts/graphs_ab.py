"""CUDA-graph replay of the blocked Cholesky against kernel-by-kernel launches: identical results, timings of a single
energy, of energy + gradient and of a 24-candidate batch at several sizes.  usage: graphs_ab.py [sizes]"""
import sys, time, json, math
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
sizes = [int(s) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["4096", "8192", "16384", "32768"])]
for n in sizes:
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 8)); w = rng.standard_normal(8) / math.sqrt(8)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(math.sqrt(8), 1.0)._spec(noise=0.1)
    specs = [dg.SEKernel(math.sqrt(8) * (1 + 0.02 * i), 1.0)._spec(noise=0.1 + 0.01 * i) for i in range(24 if n <= 16384 else 6)]
    rec, ref = {"n": n}, {}
    for graphs in (0, 1):
        ctx.set_option("graphs", graphs)
        for _ in range(3):
            e = ctx.energy(data, spec)
        reps = 20 if n <= 8192 else (8 if n <= 16384 else 3)
        t0 = time.perf_counter()
        for _ in range(reps):
            e = ctx.energy(data, spec)
        te = (time.perf_counter() - t0) / reps
        potrf_ms = ctx.timings()["potrf"]
        for _ in range(2):
            eg, g = ctx.energy_grad(data, spec)
        t0 = time.perf_counter()
        for _ in range(reps):
            eg, g = ctx.energy_grad(data, spec)
        tg = (time.perf_counter() - t0) / reps
        ctx.energy_batch(data, specs)
        t0 = time.perf_counter()
        eb, _ = ctx.energy_batch(data, specs)
        tb = time.perf_counter() - t0
        ref[graphs] = (e, eg, tuple(g), tuple(eb))
        rec[f"graphs={graphs}"] = {"energy_ms": round(te * 1e3, 3), "potrf_ms": round(potrf_ms, 3), "energy_grad_ms": round(tg * 1e3, 3),
                                   "evals_per_s": round(1 / tg, 2), "batch_cand_per_s": round(len(specs) / tb, 2)}
    rec["identical"] = ref[0] == ref[1]
    print(json.dumps(rec), flush=True)
    ctx.data_destroy(data); ctx.trim()
ctx.set_option("graphs", 1)
