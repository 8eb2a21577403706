"""Assembly workload for ncu / timing: lower tiles of K + noise for the bench configurations."""
import sys, math
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

ctx = dg.default_context()
for (n, d, name, k) in [(16384, 16, "matern52", dg.GenericMaternKernel(4.0, 2)), (16384, 8, "se", dg.SEKernel(math.sqrt(8), 1.0)),
                        (16384, 8, "ard", dg.MahalanobisKernel([3.0] * 8, 1.0))]:
    X, y, _ = orc.synthetic(n, d, 0)
    data = ctx.data_create(X, y)
    spec = k._spec(noise=0.1)
    ctx.time_phase(data, spec, 0, 1)
    ms = ctx.time_phase(data, spec, 0, 5)
    print(f"assemble {name} n={n} d={d}: {ms:.3f} ms  {8*n*(n+1)/2/ms/1e6:.1f} GB/s lower-triangle bytes", flush=True)
    ctx.data_destroy(data)
