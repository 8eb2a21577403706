"""The two other GEMM shapes of the gradient path in isolation, for ncu: K^-1 = U U^T (NT, lower tiles, k >= row; phase 3)
and an NN product of the triangular inverse (phase 4), as bench.py times them for roofline.kernels.
usage: prof_inverse.py [N]"""
import sys, math
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = 16
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); y = rng.standard_normal(n)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(math.sqrt(d), 2)._spec(noise=0.1)
ms3 = ctx.time_phase(data, spec, 3, reps=1)
ms4 = ctx.time_phase(data, spec, 4, reps=1)
m = min(n, 8192)
print(f"lauum n={n}: {ms3:.2f} ms {n**3/3/ms3/1e9:.2f} TF/s;  NN {m}^3: {ms4:.2f} ms {2*m**3/ms4/1e9:.2f} TF/s")
