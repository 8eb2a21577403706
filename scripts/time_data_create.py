"""Host-side cost of dgp_data_create / dgp_data_destroy at the bench workload (part of the e2e number)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import dynaml_b200 as dg
ctx = dg.default_context()
rng = np.random.default_rng(0)
X = rng.standard_normal((32768, 16)); y = rng.standard_normal(32768)
h = ctx.data_create(X, y); ctx.data_destroy(h)
tc = td = 0.0
for _ in range(10):
    t0 = time.perf_counter(); h = ctx.data_create(X, y); t1 = time.perf_counter(); ctx.data_destroy(h); t2 = time.perf_counter()
    tc += t1 - t0; td += t2 - t1
print(f"data_create {tc*100:.2f} ms  data_destroy {td*100:.2f} ms")
