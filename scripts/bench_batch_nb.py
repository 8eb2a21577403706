"""Batched energies (the C5 / CSA regime): candidates/s by outer Cholesky block and by candidates in flight.
With several candidates in flight the leaf chain of one hides under the GEMMs of the others, so the block width can
favour GEMM efficiency.  usage: bench_batch_nb.py [N] [candidates]"""
import sys, math, time, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ncand = int(sys.argv[2]) if len(sys.argv) > 2 else 24
d = 8
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); w = rng.standard_normal(d) / math.sqrt(d)
y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
ctx = dg.default_context()
data = ctx.data_create(X, y)
specs = [dg.SEKernel(math.sqrt(d) * (1.5 + 0.05 * i), 1.0)._spec(noise=0.2 + 0.01 * i) for i in range(ncand)]
for streams in (3, 4):
    ctx.set_option("batch_streams", streams)
    row = {}
    for nb in (0, 256, 384, 512, 768):
        ctx.set_option("nb", nb)
        ctx.energy_batch(data, specs[:streams])
        t0 = time.perf_counter(); e = ctx.energy_batch(data, specs); dt = time.perf_counter() - t0
        row["auto" if nb == 0 else nb] = round(ncand / dt, 2)
    print(json.dumps({"n": n, "candidates": ncand, "batch_streams": streams, "candidates_per_s_by_nb": row}), flush=True)
ctx.set_option("nb", 0); ctx.set_option("batch_streams", 0)
