"""Config C5: GridSearch / CSA-style landscape over 256 hyper-parameter candidates at N = 16384, candidates
sharded round-robin across the ranks of one box (run under torchrun; works on one GPU too).
Prints one JSON line on rank 0: candidates/s over the whole job."""
import json, math, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n = int(os.environ.get("C5_N", 16384))
ncand_side = int(os.environ.get("C5_GRID", 16))     # 16 x 16 = 256 candidates over (bandwidth, noiseLevel)
d = 8
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)); w = rng.standard_normal(d) / math.sqrt(d)
y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
se = dg.SEKernel(2.0 * math.sqrt(d), 1.0)
se.block("amplitude")
model = dg.GPRegression(se, dg.DiracKernel(0.5), list(zip(X, y)), ctx=dg.Context(local))
gs = dg.GridSearch(model).setGridSize(ncand_side).setStepSize(0.08).setLogScale(True)
init = {"bandwidth": 2.5 * math.sqrt(d), "noiseLevel": 0.6}
warm = gs.getGrid(init)[: 3 * world]
from dynaml_b200.optimization import sharded_energies
sharded_energies(model, warm)                      # warm-up: allocations, tile tables
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
land = gs.getEnergyLandscape(init)
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
best = min(land, key=lambda t: t[0])
if rank == 0:
    print(json.dumps({"config": "C5", "n": n, "candidates": len(land), "n_gpus": world, "seconds": dt,
                      "candidates_per_s": len(land) / dt, "tflops": len(land) * n ** 3 / 3 / dt / 1e12,
                      "best_energy": best[0], "best": best[1]}), flush=True)
if world > 1:
    dist.destroy_process_group()
