"""The dominant launch of the bench step in isolation, for ncu: the first trailing update of the C3
Cholesky (M = N - nb rows, K = nb = 768 at this N, lower tiles) -- the same launch bench.py times for `roofline`."""
import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

ctx = dg.default_context()
n, d = 32768, 16
X, y, _ = orc.synthetic(n, d, 0)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(math.sqrt(d), 2)._spec(noise=0.1)
nb = ctx.outer_block(n)
ms = ctx.time_phase(data, spec, 2, reps=3)
mt = (n - nb) // 128
print(f"trailing update n={n} K={nb}: {ms:.3f} ms  {2.0*nb*128*128*(mt*(mt+1)//2)/ms/1e9:.2f} TF/s")
