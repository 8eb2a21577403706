"""Stage hand-over of the bulk-copy GEMM: correctness soak and cost of each mode of option
`gemm_handover` (0 address dependency on the fragment registers, 1 producer-side fence.proxy.async, 2 consumer-side
fence.proxy.async, 3 plain arrival).  Every energy is compared bit for bit with the cp.async kernel's; the C-tile
prefetch is switched off on alternate repetitions (the condition that exposed the hazard).
usage: handover_modes.py [reps at N = 8192] [modes, e.g. 0123]"""
import sys, time, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
modes = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "0123")]


def synthetic(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)); w = rng.standard_normal(d) / np.sqrt(d)
    return X, np.sin(X @ w) + 0.1 * rng.standard_normal(n)


X, y = synthetic(8192, 8, 8192)
data = ctx.data_create(X, y)
spec = dg.SEKernel(float(np.sqrt(8)), 1.0)._spec(noise=0.05)
ctx.set_option("gemm_variant", 2)
e_ref = ctx.energy(data, spec)
eg_ref, g_ref = ctx.energy_grad(data, spec)
ctx.set_option("gemm_variant", 4)
Xb, yb = synthetic(32768, 16, 0)
big = ctx.data_create(Xb, yb)
bspec = dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)
out = []
for mode in modes:
    ctx.set_option("gemm_handover", mode)
    bad = 0
    t0 = time.time()
    for rep in range(reps):
        ctx.set_option("gemm_no_prefetch", rep % 2)
        bad += int(ctx.energy(data, spec) != e_ref)
        if rep % 10 == 0:
            eg, g = ctx.energy_grad(data, spec)
            bad += int(eg != eg_ref or not np.array_equal(np.asarray(g), np.asarray(g_ref)))
    ctx.set_option("gemm_no_prefetch", 0)
    soak_s = time.time() - t0
    potrf8 = ctx.time_phase(data, spec, 1, reps=5)
    potrf32 = ctx.time_phase(big, bspec, 1, reps=3)
    gemm32 = ctx.time_phase(big, bspec, 2, reps=5)
    rec = {"mode": mode, "evaluations": reps + (reps + 9) // 10, "mismatches": bad, "soak_s": round(soak_s, 1),
           "potrf_8192_ms": round(potrf8, 3), "potrf_32768_ms": round(potrf32, 2), "trailing_gemm_32768_ms": round(gemm32, 3)}
    out.append(rec)
    print(json.dumps(rec), flush=True)
ctx.set_option("gemm_handover", 0)
