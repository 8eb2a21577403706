"""Outer block of the blocked Cholesky at mid N (after the tile leaf): potrf time per nb.  usage: tune_nb2.py"""
import sys, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
ctx = dg.default_context()
for n in (4096, 8192, 12288, 16384, 24576):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 8)); y = rng.standard_normal(n)
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(2.8, 1.0)._spec(noise=0.1)
    out = {}
    for nb in (128, 256, 384, 512, 768):
        ctx.set_option("nb", nb)
        out[nb] = round(ctx.time_phase(data, spec, 1, reps=5), 3)
    ctx.set_option("nb", 0)
    print(json.dumps({"n": n, "potrf_ms_by_nb": out, "auto_nb": ctx.outer_block(n)}), flush=True)
    ctx.data_destroy(data); ctx.trim()
