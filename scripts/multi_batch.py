"""Candidates over every visible GPU from ONE process (dgp_energy_batch_multi) against one GPU, with and without the
CUDA-graph replay of the Cholesky.  usage: multi_batch.py [n] [candidates]"""
import sys, time, json, math
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from dynaml_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ncand = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ndev = _capi.device_count()
rng = np.random.default_rng(0)
X = rng.standard_normal((n, 8)); w = rng.standard_normal(8) / math.sqrt(8)
y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
se = dg.SEKernel(2.0 * math.sqrt(8), 1.0)
se.block("amplitude")
model = dg.GPRegression(se, dg.DiracKernel(0.5), list(zip(X, y)))
cands = [{"bandwidth": 2.0 * math.sqrt(8) * (1 + 0.01 * i), "noiseLevel": 0.3 + 0.005 * i} for i in range(ncand)]
out = {"n": n, "candidates": ncand, "devices": ndev}
ref = None
for devs in ([0], list(range(ndev))):
    model.useDevices(devs)
    for graphs in (0, 1):
        for c in [model.ctx] + [c for c, _ in model._peers]:
            c.set_option("graphs", graphs)
        model.energyBatch(cands[: 3 * len(devs)])
        model.energyBatch(cands[: 3 * len(devs)])
        t0 = time.perf_counter()
        e = model.energyBatch(cands)
        dt = time.perf_counter() - t0
        ref = ref or e
        out[f"devices={len(devs)},graphs={graphs}"] = {"cand_per_s": round(ncand / dt, 2), "identical": e == ref}
print(json.dumps(out))
