import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc
ctx = dg.default_context()
n, d = 16384, 16
X, y, _ = orc.synthetic(n, d, 0)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)
for mode in (2, 1, 2, 1):
    ctx.set_option("assembly", mode)
    ctx.time_phase(data, spec, 0, 1)
    ms = ctx.time_phase(data, spec, 0, 5)
    print(f"assembly mode {mode}: {ms:.3f} ms {8*n*(n+1)/2/ms/1e6:.1f} GB/s", flush=True)
