"""Config C5: batched energies for a GridSearch candidate list (N = 16384), candidates/s vs batch_streams."""
import sys, math, time, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

n, d, ncand = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 8, int(sys.argv[2]) if len(sys.argv) > 2 else 32
ctx = dg.default_context()
X, y, _ = orc.synthetic(n, d, 0)
model = dg.GPRegression(dg.SEKernel(math.sqrt(d), 1.0), dg.DiracKernel(0.1), list(zip(X, y)))
gs = dg.GridSearch(model).setGridSize(int(round(ncand ** (1 / 3)))).setStepSize(0.15).setLogScale(True)
grid = gs.getGrid({"bandwidth": 2 * math.sqrt(d), "amplitude": 1.5, "noiseLevel": 0.3})[:ncand]
out = {}
ref = None
for streams in (1, 2, 3):
    ctx.set_option("batch_streams", streams)
    model.energyBatch(grid[:streams * 2])
    t0 = time.perf_counter(); e = model.energyBatch(grid); dt = time.perf_counter() - t0
    if ref is None: ref = e
    assert e == ref, "concurrent batches must be bitwise identical to the serial one"
    print(f"n={n} candidates={len(grid)} batch_streams={streams}: {dt*1e3:.1f} ms  {len(grid)/dt:.2f} cand/s  {len(grid)*n**3/3/dt/1e12:.2f} TF/s", flush=True)
    out[streams] = len(grid) / dt
json.dump(out, open(f"gpurun_out/batch_{n}.json", "w"))
