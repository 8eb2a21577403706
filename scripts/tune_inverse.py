"""K^-1 for the gradient: W = L^-1 then W^T W (NN, NN, TN products) against U = L^-T then U U^T (NT, NN, NT products):
option inverse_transposed.  Whole C3 step, and the agreement of the two gradients."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

ctx = dg.default_context()
for (n, d) in ((8192, 8), (32768, 16)):
    X, y, _ = orc.synthetic(n, d, 0)
    data = ctx.data_create(X, y)
    spec = dg.GenericMaternKernel(float(np.sqrt(d)), 2)._spec(noise=0.1)
    out = {}
    for mode in (0, 1, 0, 1):
        ctx.set_option("inverse_transposed", mode)
        ctx.energy_grad(data, spec)
        e, g = ctx.energy_grad(data, spec)
        tm = ctx.timings()
        out[mode] = (e, np.array(g[1:], dtype=float))
        print(f"n={n} inverse_transposed={mode}: total {tm['total']:.1f} ms potrf {tm['potrf']:.1f} inverse {tm['inverse']:.1f}"
              f" E={e!r} g={[float(v) for v in g[1:]]!r}", flush=True)
    print("  relative gradient difference between the two:", float(np.abs(out[0][1] - out[1][1]).max() / np.abs(out[0][1]).max()), flush=True)
    ctx.data_destroy(data)
