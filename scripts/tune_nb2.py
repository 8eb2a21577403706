"""Outer block of the blocked Cholesky: potrf time per fixed nb and with the automatic choice (nb = 0: the block
width follows the size of the remaining trailing matrix).  usage: tune_nb2.py [sizes, comma separated]"""
import sys, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
ctx = dg.default_context()
sizes = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [4096, 8192, 12288, 16384, 24576, 32768]
for n in sizes:
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 8)); y = rng.standard_normal(n)
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(2.8, 1.0)._spec(noise=0.1)
    out = {}
    for nb in (256, 512, 768):
        ctx.set_option("nb", nb)
        out[nb] = round(ctx.time_phase(data, spec, 1, reps=5), 3)
    ctx.set_option("nb", 0)
    out["auto"] = round(ctx.time_phase(data, spec, 1, reps=5), 3)
    print(json.dumps({"n": n, "potrf_ms_by_nb": out, "auto_first_block": ctx.outer_block(n)}), flush=True)
    ctx.data_destroy(data); ctx.trim()
