import sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
import dynaml_b200 as dg
ctx = dg.default_context()
n = 32768
rng = np.random.default_rng(0)
X = rng.standard_normal((n, 16)); y = rng.standard_normal(n)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)
out = {}
for nb in (512, 640, 768, 896, 1024):
    ctx.set_option("nb", nb)
    out[nb] = round(ctx.time_phase(data, spec, 1, reps=3), 2)
print(json.dumps({"n": n, "potrf_ms_by_nb": out}))
