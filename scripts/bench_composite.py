"""Cost of the generic composite path next to the specialised single-kernel path (N = 16384, d = 8)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import dynaml_b200 as dg

ctx = dg.default_context()
ctx.set_option("timing", 1)
n, d = 16384, 8
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
data = ctx.data_create(X, y)
cases = {
    "se (single-kernel path)": dg.SEKernel(3.0, 1.0),
    "se + matern52": dg.SEKernel(3.0, 1.0) + dg.GenericMaternKernel(4.0, 2),
    "se * periodic + rbf*0.5": dg.SEKernel(3.0, 1.0) * dg.PeriodicKernel(1.0, 2.0, 0.1) + dg.RBFKernel(5.0) * 0.5,
    "committee of 9 se": dg.DecomposableCovariance([dg.SEKernel(2.0 + 0.3 * i, 1.0) for i in range(9)], [1.0 / 9] * 9),
}
for name, k in cases.items():
    spec = k._spec(noise=0.2)
    ctx.energy_grad(data, spec)
    ctx.energy_grad(data, spec)
    t = ctx.timings()
    print(f"{name:28s} total {t['total']:8.2f} ms  assemble {t['assemble']:7.2f}  potrf {t['potrf']:7.2f}  "
          f"inverse {t['inverse']:7.2f}  trace {t['trace']:7.2f}", flush=True)
ctx.data_destroy(data)
