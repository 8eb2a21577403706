import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import dynaml_b200 as dg
ctx = dg.default_context()
n = 32768
for d, mk in ((8, lambda: dg.SEKernel(2.0, 1.0)._spec(noise=0.1)), (16, lambda: dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1))):
    rng = np.random.default_rng(0)
    data = ctx.data_create(rng.standard_normal((n, d)), rng.standard_normal(n))
    for dbg in (None, "7", "8", None):
        if dbg is None: os.environ.pop("DGP_LEAN_DBG", None)
        else: os.environ["DGP_LEAN_DBG"] = dbg
        ctx.time_phase(data, mk(), 0, 2)
        print(d, dbg, ctx.time_phase(data, mk(), 0, 10), flush=True)
    ctx.data_destroy(data)
