"""One energy (or energy + gradient) evaluation at a given size, for launch lists (ncu --metrics gpu__time_duration.sum).
usage: one_energy.py N [grad] [kernel: ard|se|matern]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
grad = len(sys.argv) > 2 and sys.argv[2] == "grad"
kern = sys.argv[3] if len(sys.argv) > 3 else "ard"
d = 16 if kern == "matern" else 8
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); w = rng.standard_normal(d) / np.sqrt(d)
y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
ctx = dg.default_context()
data = ctx.data_create(X, y)
k = {"ard": dg.MahalanobisKernel([float(np.sqrt(d))] * d, 1.0), "se": dg.SEKernel(float(np.sqrt(d)), 1.0),
     "matern": dg.GenericMaternKernel(float(np.sqrt(d)), 2)}[kern]
spec = k._spec(noise=0.1)
print(ctx.energy_grad(data, spec) if grad else ctx.energy(data, spec), ctx.timings())
