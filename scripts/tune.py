"""Tuning sweep on a B200: outer block size / look-ahead for the blocked Cholesky."""
import sys, math, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

ctx = dg.default_context()
res = {}
for (n, d) in [(8192, 8), (16384, 8), (32768, 16)]:
    X, y, _ = orc.synthetic(n, d, 0)
    data = ctx.data_create(X, y)
    spec = dg.GenericMaternKernel(math.sqrt(d), 2)._spec(noise=0.1)
    for nb in (256, 384, 512, 768, 1024):
        for la in ((1, 0) if len(sys.argv) > 1 else (1,)):  # any argument: also time look-ahead off
            ctx.set_option("nb", nb); ctx.set_option("lookahead", la)
            ctx.time_phase(data, spec, 1, 1)
            ms = ctx.time_phase(data, spec, 1, 3)
            print(f"potrf n={n} nb={nb} la={la}: {ms:.2f} ms {n**3/3/ms/1e9:.2f} TF/s", flush=True)
            res[f"{n}_{nb}_{la}"] = ms
    ctx.set_option("nb", 256); ctx.set_option("lookahead", 1)
    ctx.data_destroy(data)
json.dump(res, open("gpurun_out/tune_nb.json", "w"), indent=1)
