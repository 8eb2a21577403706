"""Distributed 2D block-cyclic Cholesky + NLL on the GPUs of one box (run under torchrun).

    torchrun --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 scripts/dist_run.py \
        --n 65536 --tile 512 [--check 4096]

Prints one JSON line on rank 0: factorisation TFLOP/s = (N^3/3) / max-over-ranks device time."""
import argparse, json, math, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
import dynaml_b200 as dg
from oracle import gp_oracle as orc

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}

ap = argparse.ArgumentParser()
ap.add_argument("--gp-n", dest="n", type=int, default=65536)
ap.add_argument("--gp-d", dest="d", type=int, default=8)
ap.add_argument("--gp-tile", dest="tile", type=int, default=512)
ap.add_argument("--gp-check", dest="check", type=int, default=3000, help="size of the oracle parity check run first (0 = skip)")
ap.add_argument("--gp-reps", dest="reps", type=int, default=2)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = dg.Context(local)
pr, pc = GRIDS[world]
uid = [ctx.dist_unique_id() if (rank == 0 and world > 1) else None]
if world > 1:
    dist.broadcast_object_list(uid, src=0)
ctx.dist_init(uid[0], rank, world, pr, pc)

out = {"world": world, "grid": [pr, pc], "tile": args.tile}
if args.check:
    X, y, _ = orc.synthetic(args.check, args.d, seed=5)
    data = ctx.data_create(X, y)
    spec = dg.SEKernel(math.sqrt(args.d), 1.0)._spec(noise=0.1)
    e, _ = ctx.dist_energy(data, spec, 256)
    ctx.data_destroy(data)
    if rank == 0:
        eo = orc.energy(orc.Kernel("se", {"bandwidth": math.sqrt(args.d), "amplitude": 1.0}), 0.1, X, y)
        out["check"] = {"n": args.check, "energy": e, "oracle": eo, "rel": abs(e - eo) / abs(eo)}
        assert abs(e - eo) <= 1e-9 * abs(eo), out["check"]

X, y, _ = orc.synthetic(args.n, args.d, seed=0)
data = ctx.data_create(X, y)
spec = dg.SEKernel(math.sqrt(args.d), 1.0)._spec(noise=0.1)
times, energies = [], []
for rep in range(args.reps + 1):
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e, ms = ctx.dist_energy(data, spec, args.tile)
    wall = time.perf_counter() - t0
    t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rep > 0:
        times.append(t.tolist())
    energies.append(e)
ctx.data_destroy(data)
ctx.dist_finalize()
if rank == 0:
    ms = min(t[0] for t in times)
    out.update({"n": args.n, "energy": energies[-1], "chol_ms": ms, "wall_ms": min(t[1] for t in times),
                "chol_tflops": args.n ** 3 / 3 / (ms * 1e-3) / 1e12, "per_gpu_tflops": args.n ** 3 / 3 / (ms * 1e-3) / 1e12 / world})
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
