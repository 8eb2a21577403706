"""Workload for ncu captures of the two lean Gram kernels (assembly and gradient trace pass): Matern-5/2 d = 16 and
SE d = 8 at N = 16384, one energy + gradient evaluation each.
    ncu --set full --import-source on -k regex:"assemble_gram_lean|grad_trace_lean" python scripts/prof_lean.py"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for d, k in ((16, dg.GenericMaternKernel(4.0, 2)), (8, dg.SEKernel(float(np.sqrt(8)), 1.0))):
    rng = np.random.default_rng(d)
    X = rng.standard_normal((n, d)); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    data = ctx.data_create(X, y)
    e, g = ctx.energy_grad(data, k._spec(noise=0.1))
    print(d, e, list(g), ctx.timings()["assemble"], ctx.timings()["trace"])
    ctx.data_destroy(data)
