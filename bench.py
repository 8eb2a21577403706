#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native DynaML GP hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): "GP energy+grad evals/s (N=32k FP64); dist. Cholesky FP64 TFLOP/s at 1-8 GPUs".

`value` / `e2e` -- a step is ONE evaluation of `energy` + `gradEnergy` (dgp_energy_grad: kernel-matrix assembly ->
blocked Cholesky -> solves -> K^-1 -> gradient traces) at config C3 -- synthetic N = 32768, d = 16, Matern-5/2 + Dirac
noise -- with X, y already resident in HBM (`value`), and the same evaluation through the C ABI from host buffers
including the H2D of X, y and the D2H of the results (`e2e`).  At N GPUs every rank evaluates its own hyper-parameter
candidate per step (candidates are independent; no data-path collective), so `value` = N*K evaluations /
max-over-ranks time and scaling is "weak".

The two sharded configurations of the north star run in the same invocation, at every N, and are reported under
`roofline` (the driver keeps that object verbatim):
  roofline.dist_c4 -- config C4: dgp_dist_energy at N = 131072, d = 8, SE, tile 512 on the 1x1 / 1x2 / 2x2 / 2x4 process
                      grid: 2D block-cyclic Cholesky + NLL over the library's own NCCL communicator; device time of the
                      factorisation as the max over ranks, N^3/3 flop / time, per-GPU fraction of the DMMA peak, the energy,
                      and a parity leg at n = 8192 on the same grid against the dense one-GPU path and the committed oracle
                      value (tests/golden/large.json);
  roofline.c5      -- config C5: 256 hyper-parameter candidates at N = 16384 through GridSearch.getEnergyLandscape,
                      candidates sharded round-robin over the ranks -> candidates/s, with three candidates compared
                      against committed oracle energies;
  roofline.predict -- (rank 0) the prediction half of config C3: fit, then mean + covariance + error
                      bars of 4096 test points: device times of the triangular solve and of V'V with their fractions of the
                      DMMA peak, the wall time of dgp_predict including the copy-out, size-independent checks.

Roofline: the dominant kernel is the FP64 DMMA GEMM (dgemm_dmma_v4_kernel).  `achieved` is the algorithmic flops of
the launch the Cholesky repeats -- the first trailing update, C -= P P^T over the lower tiles with M = N - nb rows and
K = nb = 768 at this N (2 * K * 128^2 flop per lower tile) -- divided by that launch's CUDA-event duration, measured
live after the timed steps on the library's own stream; `peak` is the FP64 DMMA peak measured live by a
register-resident mma.sync.m8n8k4.f64 loop (MEASURED_PEAKS.json carries no FP64 figure).
`roofline.kernels` gives the same fraction per GEMM instantiation; `step_frac` is the whole step: N^3 flop / step
time / peak.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "GP energy+grad evals/s (N=32k FP64)"
UNIT = "evals/s"
WORKLOADS = {
    # name: (N, d, kernel)
    "c3": (32768, 16, "matern"),   # the configuration the metric is quoted on
    "c2": (8192, 8, "ard"),
    "c5": (16384, 8, "se"),
    "small": (2048, 8, "matern"),  # smoke-sized, for quick checks
}
GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-n", type=int, default=4096, help="size the O(N^2) part of the CPU arm is timed at")
    ap.add_argument("--cpu-sample-n3", type=int, default=16384, help="size the O(N^3) LAPACK part is timed at")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4-n", type=int, default=131072, help="size of the distributed-Cholesky leg (0 = skip)")
    ap.add_argument("--c4-tile", type=int, default=512)
    ap.add_argument("--c5-n", type=int, default=16384, help="size of the GridSearch leg (0 = skip)")
    ap.add_argument("--c5-grid", type=int, default=16, help="grid points per hyper-parameter (16 x 16 = 256 candidates)")
    ap.add_argument("--predict-m", type=int, default=4096, help="test points of the prediction leg (0 = skip)")
    return ap.parse_args()


def workload_config(name: str) -> dict:
    n, d, kern = WORKLOADS[name]
    desc = {"matern": "Matern-5/2 (GenericMaternKernel p=2)", "ard": "ARD-SE (MahalanobisKernel)", "se": "SEKernel"}[kern]
    return {"workload": f"synthetic N={n}, d={d}, {desc} + DiracKernel noise: energy + gradEnergy per step",
            "N": n, "d": d, "kernel": kern, "data_seed": 0,
            "l2": "per-step working set (K, L^-1, K^-1: 3 x 8*N^2 bytes) exceeds the 126 MB L2; no flush needed"}


def synthetic(n: int, d: int, seed: int = 0):
    """SURVEY 8(d) inputs: X ~ N(0,1)^{n x d}; y = sin(X w) + 0.1 eps, w ~ N(0, I/d).  (Same generator as
    oracle.gp_oracle.synthetic; restated here so the measured arm does not touch oracle/.)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    w = rng.standard_normal(d) / math.sqrt(d)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    return X, y


def make_kernel(kern: str, d: int, rank: int):
    """Device kernel object; each rank gets its own candidate (length-scale scaled by the rank)."""
    import dynaml_b200 as dg
    scale = 1.0 + 0.05 * rank
    if kern == "matern":
        return dg.GenericMaternKernel(math.sqrt(d) * scale, 2)
    if kern == "ard":
        return dg.MahalanobisKernel([math.sqrt(d) * scale] * d, 1.0)
    return dg.SEKernel(math.sqrt(d) * scale, 1.0)


def golden_large() -> dict:
    """Committed oracle values at the BASELINE sizes (tests/golden/large.json, made by make_golden_large.py)."""
    try:
        return json.loads((ROOT / "tests" / "golden" / "large.json").read_text())
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU arm ---------------------------------------------------------------------------------------------------
def _all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core at every N.  Set the
    thread-count variables before NumPy / SciPy load their BLAS."""
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except Exception:
        n = max(1, os.cpu_count() or 1)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)
    return n


def cpu_sample(name: str, n2: int, n3: int) -> dict:
    """One sample of the oracle port on the host cores, scaled part by part to the workload's N."""
    from oracle import cpu_baseline as cb
    n, d, kern = WORKLOADS[name]
    n2, n3 = min(n2, n), min(max(n2, n3), n)
    cb.use_all_cores()
    s = cb.time_parts(n2, n3, d, kern)
    sec_full = cb.extrapolate(s["t2"], s["t3"], s["n2"], n, s["n3"])
    s.update({"sec_full": sec_full, "N": n, "d": d, "kern": kern})
    return s


def cpu_record(samples: list) -> dict:
    s0 = samples[0]
    sec = sum(s["sec_full"] for s in samples) / len(samples)
    t2 = sum(s["t2"] for s in samples) / len(samples)
    t3 = sum(s["t3"] for s in samples) / len(samples)
    extrap = s0["n2"] < s0["N"] or s0["n3"] < s0["N"]
    return {"value": 1.0 / sec, "unit": UNIT, "cores": s0["threads"], "kind": "port", "host_cpus": os.cpu_count(),
            "extrapolated": extrap, "sample_n": s0["n3"], "sample_n2": s0["n2"], "sample_seconds": t2 + t3,
            "sample": (f"one energy+grad of the same workload (d={s0['d']}, {s0['kern']}) with NumPy/SciPy-OpenBLAS on "
                       f"{s0['threads']} threads: O(N^2) part (direct-difference assembly, gradient traces) {t2:.2f} s at "
                       f"N={s0['n2']}; O(N^3) part (dpotrf, 2 x dtrtrs, dpotri) {t3:.2f} s at N={s0['n3']}; "
                       + (f"EXTRAPOLATED to N={s0['N']} as t2*(N/n2)^2 + t3*(N/n3)^3 = {sec:.1f} s per evaluation"
                          if extrap else f"measured at the full N={s0['N']}: {sec:.1f} s per evaluation")
                       + "; the JVM reference cannot run in this image")}


def run_reference(args) -> None:
    """--impl reference: DynaML is JVM/Scala on Breeze + netlib-java; no JVM, sbt or jars exist in
    this image, so the reference arm is the oracle port on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _all_host_threads()
    from oracle import cpu_baseline as cb
    n, d, kern = WORKLOADS[args.workload]
    cb.use_all_cores()
    for _ in range(max(0, min(args.warmup, 1))):
        cb.time_parts(min(1024, n), min(2048, n), d, kern)
    samples = [cpu_sample(args.workload, args.cpu_sample_n, args.cpu_sample_n3) for _ in range(max(1, args.steps))]
    rec = cpu_record(samples)
    sec = 1.0 / rec["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload), "extrapolated": rec["extrapolated"],
        "cpu_baseline": rec,
        "e2e": {"value": rec["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- sharded configurations ------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing of the multi-rank run (barrier, max over ranks, object broadcast); no-ops at N = 1."""

    def __init__(self, world: int, local: int):
        self.world, self.local, self.pg = world, local, None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
            self.pg = dist
            # host-side barrier (sockets): ranks waiting in an NCCL barrier keep a spinning kernel on their GPU, which the
            # driver time-slices against any other process computing on that GPU (the single-process C5 leg)
            self.host = dist.new_group(backend="gloo")

    def barrier(self):
        if self.pg is not None:
            self.pg.barrier()

    def host_barrier(self):
        if self.pg is not None:
            import torch
            torch.cuda.synchronize()
            self.pg.barrier(group=self.host)

    def max(self, x: float) -> float:
        if self.pg is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{self.local}")
        self.pg.all_reduce(t, op=self.pg.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj) -> list:
        if self.pg is None:
            return [obj]
        out = [None] * self.world
        self.pg.all_gather_object(out, obj)
        return out

    def bcast(self, obj):
        if self.pg is None:
            return obj
        box = [obj]
        self.pg.broadcast_object_list(box, src=0)
        return box[0]

    def close(self):
        if self.pg is not None:
            self.pg.destroy_process_group()


def run_c4(ctx, D: Dist, rank: int, args, dmma_peak: float | None) -> dict:
    """Config C4: 2D block-cyclic distributed Cholesky + NLL (dgp_dist_energy) on this run's process grid."""
    import dynaml_b200 as dg
    world = D.world
    pr, pc = GRIDS[world]
    uid = D.bcast(ctx.dist_unique_id() if (rank == 0 and world > 1) else None)
    ctx.dist_init(uid, rank, world, pr, pc)
    info = ctx.dist_info()
    gold = golden_large().get("c4_check", {})
    spec = dg.SEKernel(math.sqrt(8), 1.0)._spec(noise=0.1)
    from dynaml_b200._capi import dist_plan
    plan = dist_plan(args.c4_n, args.c4_tile, pr, pc, rank)
    out = {"config": f"C4: synthetic N={args.c4_n}, d=8, SEKernel + DiracKernel(0.1), tile {args.c4_tile}, "
                     f"{pr}x{pc} process grid", "n": args.c4_n, "tile": args.c4_tile, "grid": [pr, pc], "n_gpus": world,
           "local_matrix_gb_rank0": plan["local_bytes"] / 1e9, "owned_lower_tiles_rank0": plan["owned_lower_tiles"],
           "library_nccl_ranks": D.gather(info["nccl_ranks"]), "nccl_version": info["nccl_version"]}
    # -- parity leg: same grid and communicator at n = 8192 against the dense one-GPU path and the committed oracle value
    nchk = int(gold.get("n", 8192))
    Xc, yc = synthetic(nchk, 8, seed=0)
    dc = ctx.data_create(Xc, yc)
    e_chk, _ = ctx.dist_energy(dc, spec, args.c4_tile)
    e_dense = ctx.energy(dc, spec)
    ctx.data_destroy(dc)
    chk = {"n": nchk, "energy": e_chk, "dense_one_gpu": e_dense, "rel_vs_dense": abs(e_chk - e_dense) / abs(e_dense),
           "identical_on_all_ranks": len(set(D.gather(e_chk))) == 1}
    if "energy" in gold:
        chk["oracle"] = gold["energy"]
        chk["rel_vs_oracle"] = abs(e_chk - gold["energy"]) / abs(gold["energy"])
    out["parity_check"] = chk
    ctx.trim()
    # -- the timed factorisation
    X, y = synthetic(args.c4_n, 8, seed=0)
    data = ctx.data_create(X, y)
    reps = 1 if world == 1 else 2
    times, energies = [], []
    for _ in range(reps):
        D.barrier()
        t0 = time.perf_counter()
        e, ms = ctx.dist_energy(data, spec, args.c4_tile)
        wall = time.perf_counter() - t0
        times.append((D.max(ms), D.max(wall * 1e3)))
        energies.append(e)
    all_e = D.gather(energies[-1])
    ctx.data_destroy(data)
    ctx.dist_finalize()
    ms = min(t[0] for t in times)
    tf = args.c4_n ** 3 / 3.0 / (ms * 1e-3) / 1e12
    out.update({"chol_ms": ms, "call_ms": min(t[1] for t in times), "reps": reps,
                "timing": "CUDA events on the library's stream around the factorisation loop, max over ranks, best of reps",
                "chol_tflops": tf, "per_gpu_tflops": tf / world, "energy": energies[-1],
                "energy_identical_on_all_ranks": len(set(all_e)) == 1, "energy_identical_across_reps": len(set(energies)) == 1})
    if dmma_peak:
        out["per_gpu_frac_of_dmma_peak"] = tf / world / dmma_peak
    ref = golden_large().get("c4_device_1gpu", {})
    if ref.get("n") == args.c4_n and "energy" in ref:
        out["energy_1gpu_recorded"] = ref["energy"]
        out["rel_vs_1gpu"] = abs(energies[-1] - ref["energy"]) / abs(ref["energy"])
    return out


def run_c5(ctx, D: Dist, rank: int, args) -> dict:
    """Config C5: GridSearch over c5_grid^2 (bandwidth, noiseLevel) candidates at N = c5_n, sharded over the ranks."""
    import dynaml_b200 as dg
    from dynaml_b200.optimization import sharded_energies
    n, d, world = args.c5_n, 8, D.world
    X, y = synthetic(n, d, seed=0)
    se = dg.SEKernel(2.0 * math.sqrt(d), 1.0)
    se.block("amplitude")
    model = dg.GPRegression(se, dg.DiracKernel(0.5), list(zip(X, y)), ctx=ctx)
    gs = dg.GridSearch(model).setGridSize(args.c5_grid).setStepSize(0.08).setLogScale(True)
    init = {"bandwidth": 2.5 * math.sqrt(d), "noiseLevel": 0.6}
    sharded_energies(model, gs.getGrid(init)[: 3 * world])   # warm-up: allocations, tile tables
    D.barrier()
    launches0 = ctx.launch_count()
    t0 = time.perf_counter()
    land = gs.getEnergyLandscape(init)
    dev_ms = ctx.timings()["batch_total"]
    D.barrier()
    dt = D.max(time.perf_counter() - t0)
    dev_ms = D.max(dev_ms)
    best = min(land, key=lambda t: t[0])
    out = {"config": f"C5: GridSearch.getEnergyLandscape over {len(land)} (bandwidth, noiseLevel) candidates, synthetic "
                     f"N={n}, d={d}, SEKernel + DiracKernel; candidates sharded round-robin over {world} rank(s)",
           "n": n, "candidates": len(land), "n_gpus": world, "seconds": dt, "candidates_per_s": len(land) / dt,
           "device_seconds_max_rank": dev_ms * 1e-3, "candidates_per_s_device": len(land) / (dev_ms * 1e-3),
           "tflops": len(land) * n ** 3 / 3.0 / dt / 1e12, "best_energy": best[0], "best": best[1],
           "gpu_launches_rank0": ctx.launch_count() - launches0}
    gold = golden_large().get("c5", {})
    if gold.get("n") == n and gold.get("gridsize") == args.c5_grid:
        rels = []
        for c in gold["candidates"]:
            e_dev, cfg = land[c["index"]]
            assert abs(cfg["bandwidth"] - c["bandwidth"]) <= 1e-12 * c["bandwidth"]
            rels.append(abs(e_dev - c["energy"]) / abs(c["energy"]))
        out["oracle_candidates"] = [c["index"] for c in gold["candidates"]]
        out["max_rel_vs_oracle"] = max(rels)
    # -- the same landscape from ONE process driving every GPU of the run (dgp_energy_batch_multi: DynaML is a single
    #    JVM); the other ranks have released their workspaces and wait at the barrier
    if world > 1:
        ctx.trim()
        D.host_barrier()
        if rank == 0:
            try:
                model.useDevices(list(range(world)))
                model.energyBatch(gs.getGrid(init)[: 3 * world])          # warm-up on every device
                t0 = time.perf_counter()
                land1 = [(e, c) for e, c in zip(model.energyBatch(gs.getGrid(init)), gs.getGrid(init))]
                dt1 = time.perf_counter() - t0
                out["single_process"] = {"what": "dgp_energy_batch_multi: one host process, one context per GPU, "
                                                 "candidates round-robin over the GPUs", "n_gpus": world, "seconds": dt1,
                                         "candidates_per_s": len(land1) / dt1,
                                         "identical_to_sharded_run": [e for e, _ in land1] == [e for e, _ in land]}
            except Exception as ex:
                out["single_process"] = {"error": f"{type(ex).__name__}: {ex}"}
            model.useDevices([])
        D.host_barrier()
    model._drop_data()
    return out


def main():
    args = parse()
    # NCCL prints "NCCL version ..." on stdout at the first communicator unless told otherwise; stdout carries the one
    # JSON line (an explicit NCCL_DEBUG from the caller is respected)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import dynaml_b200 as dg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    D = Dist(world, local)
    barrier, max_over_ranks = D.barrier, D.max

    ctx = dg.Context(local)
    n, d, kern = WORKLOADS[args.workload]
    X, y = synthetic(n, d, seed=0)
    kernel = make_kernel(kern, d, rank)
    spec = kernel._spec(noise=0.1)
    data = ctx.data_create(X, y)

    # ---- device-resident loop: `value` ------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        ctx.energy_grad(data, spec)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e, g = ctx.energy_grad(data, spec)   # synchronous: returns after the stream has drained
        dev_ms += ctx.timings()["total"]
    barrier()
    t1 = time.perf_counter()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    phase = ctx.timings()
    wall = max_over_ranks(t1 - t0)
    dev = max_over_ranks(dev_ms * 1e-3)
    value = world * args.steps / wall

    # ---- end to end through the C ABI from host buffers: `e2e` --------------------------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h = ctx.data_create(X, y)            # H2D of X (N*d) and y (N)
        e2, g2 = ctx.energy_grad(h, spec)    # D2H of E and the gradient inside the call
        ctx.data_destroy(h)
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * args.steps / e2e_wall
    assert e2 == e and all(a == b or (a != a and b != b) for a, b in zip(g2, g))

    line, peaks = None, None
    if rank == 0:
        # ---- roofline of the dominant kernel, measured live ----------------------------------------
        peaks = ctx.measure_peaks()
        dmma = peaks["dmma_tflops"]
        nb = ctx.outer_block(n)
        gemm_ms = ctx.time_phase(data, spec, 2, reps=5)
        mt = (n - nb) // 128
        flops = 2.0 * nb * 128 * 128 * (mt * (mt + 1) // 2)
        achieved = flops / (gemm_ms * 1e-3) / 1e12
        # per instantiation: the K^-1 = U U^T launch (NT with a triangular k-range, N^3/3: the largest launch of a
        # step) and an NN product (triangular inverse, prediction TRSM)
        lauum_ms = ctx.time_phase(data, spec, 3, reps=2)
        lauum_tf = n ** 3 / 3.0 / (lauum_ms * 1e-3) / 1e12
        mnn = min(n, 8192)
        nn_ms = ctx.time_phase(data, spec, 4, reps=3)
        nn_tf = 2.0 * mnn ** 3 / (nn_ms * 1e-3) / 1e12
        traffic = None
        rf = ROOT / "profiles" / "roofline_traffic.json"
        if rf.exists():
            try:
                traffic = json.loads(rf.read_text()).get(f"{args.workload}_nb{nb}", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        potrf_tf = (n ** 3 / 3.0) / (phase["potrf"] * 1e-3) / 1e12
        inv_tf = (2.0 * n ** 3 / 3.0) / (phase["inverse"] * 1e-3) / 1e12 if phase["inverse"] > 0 else None
        asm_gbs = 8.0 * n * (n + 1) / 2 / (phase["assemble"] * 1e-3) / 1e9
        # ---- assembly against HBM bandwidth (north star: >= 70 %), timed alone: this workload's kernel and the
        #      SE d = 8 kernel of config C4 at the same N.  Algorithmic bytes = 8 N (N + 1) / 2 (lower triangle).
        hbm_peak, hbm_src = peaks["hbm_write_gbs"], "HBM write bandwidth measured live by dgp_measure_peaks"
        mp = ROOT / "MEASURED_PEAKS.json"
        if mp.exists():
            try:
                hbm_peak, hbm_src = float(json.loads(mp.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
            except Exception:
                pass
        asm_bytes = 8.0 * n * (n + 1) / 2
        asm_ms = ctx.time_phase(data, spec, 0, reps=10)
        rng8 = np.random.default_rng(1)
        data8 = ctx.data_create(rng8.standard_normal((n, 8)), rng8.standard_normal(n))
        se_ms = ctx.time_phase(data8, dg.SEKernel(2.0, 1.0)._spec(noise=0.1), 0, reps=10)
        ctx.data_destroy(data8)
        # FP64-pipe roofline of the assembly.  On this part DMMA and DFMA-class instructions share one
        # FP64 pipe per SM sub-partition: a DFMA-class warp instruction holds it 2 clk, an m8n8k4 DMMA 16 clk (the
        # measured DMMA and DFMA peaks are both that rate).  Per 32 entries the lean kernel needs dp/8 DMMAs for the
        # Gram distance (2 dp clk) and `inst` FP64 instructions for the epilogue (2 clk each; counted in the SASS):
        # bound = entries / 32 * (2 dp + 2 inst) clk over 4 sub-partitions x SMs at the sampled SM clock.
        entries = n * (n + 1) / 2
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        sms = ctx.info()["sm_count"]
        FP64_INST = {"matern": 22, "ard": 13, "se": 13}   # exp 10-12, sqrt 6, polynomial 2-3, product, DADD seed

        def fp64_bound_ms(dp_, inst):
            return entries / 32.0 * (2.0 * dp_ + 2.0 * inst) / (sms * 4 * sm_hz) * 1e3

        wl_bound, se_bound = fp64_bound_ms(-(-d // 4) * 4, FP64_INST[kern]), fp64_bound_ms(8, FP64_INST["se"])
        roofline_assembly = {
            "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": hbm_src,
            "kernel": "assemble_gram_lean_kernel (lower tiles of K, N = %d)" % n,
            "workload_kernel": {"ms": asm_ms, "achieved": asm_bytes / asm_ms / 1e6,
                                "frac": asm_bytes / asm_ms / 1e6 / hbm_peak,
                                "fp64_bound_ms": wl_bound, "fp64_frac": wl_bound / asm_ms,
                                "binding": "fp64" if wl_bound > asm_bytes / hbm_peak / 1e6 else "hbm"},
            "se_d8": {"ms": se_ms, "achieved": asm_bytes / se_ms / 1e6, "frac": asm_bytes / se_ms / 1e6 / hbm_peak,
                      "fp64_bound_ms": se_bound, "fp64_frac": se_bound / se_ms,
                      "binding": "fp64" if se_bound > asm_bytes / hbm_peak / 1e6 else "hbm"},
            "fp64": {"what": "time the FP64 pipe alone needs: entries/32 * (2*dp + 2*inst) clk / (4 sub-partitions * SMs "
                             "* SM clock); DMMA (16 clk) and DFMA-class (2 clk) warp instructions share the pipe",
                     "inst_per_entry": FP64_INST, "sm_mhz": sm_hz / 1e6, "sms": sms},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload),
            "device_ms_per_step": dev / args.steps * 1e3,
            "roofline_assembly": roofline_assembly,
            "phases_ms": phase,
            "peaks_measured": peaks,
            "e2e": {"value": e2e_value, "unit": UNIT, # dgp_data_create copies the centred points, the raw points and the residual vector (padded) in one transfer
                    "h2d_bytes_per_step": int(8 * (2 * (-(-n // 128) * 128 + 128) * (-(-d // 4) * 4) + (-(-n // 128) * 128))),
                    "d2h_bytes_per_step": int(8 * (1 + len(g)))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # blocked hyper-parameters have no gradient (NaN in the API, as in the reference's map): null in strict JSON
            "result": {"energy": e, "grad": [None if v != v else float(v) for v in g]},
        }
        roofline = {"bound": "tensor", "kernel": f"dgemm_dmma_v4_kernel<NT> (first trailing update of the Cholesky: M = N - {nb} rows, K = {nb}, lower tiles)",
                    "achieved": achieved, "peak": dmma, "unit": "TFLOP/s",
                    "frac": achieved / dmma, "traffic": traffic,
                    "peak_source": "FP64 DMMA peak measured live by dgp_measure_peaks (register-resident "
                                   "mma.sync.m8n8k4.f64 loop); MEASURED_PEAKS.json has no FP64 entry",
                    "peak_sustained": peaks["dmma_tflops_sustained"],
                    "kernels": {
                        "nt_trailing_update": {"ms": gemm_ms, "tflops": achieved, "frac": achieved / dmma},
                        "nt_lauum_k_ge_row": {"ms": lauum_ms, "tflops": lauum_tf, "frac": lauum_tf / dmma,
                                              "what": "K^-1 = U U^T, lower tiles, k >= row: N^3/3 flop, the largest launch"},
                        "nn_full": {"ms": nn_ms, "tflops": nn_tf, "frac": nn_tf / dmma, "what": f"C = A B, M = N = K = {mnn}"}},
                    "step_tflops": n ** 3 / (wall / args.steps) / 1e12,
                    "step_frac": n ** 3 / (wall / args.steps) / 1e12 / dmma,
                    "potrf_tflops": potrf_tf, "potrf_frac": potrf_tf / dmma,
                    "inverse_tflops": inv_tf, "inverse_frac": inv_tf / dmma if inv_tf else None,
                    "assemble_gbs": asm_gbs, "assemble_frac_of_hbm_write": asm_gbs / peaks["hbm_write_gbs"]}
        line["roofline"] = roofline
        # ---- the prediction half of config C3: fit, then mean + covariance + error bars of 4096 test points --------
        if args.predict_m > 0:
            try:
                rngp = np.random.default_rng(7)
                Xs = rngp.standard_normal((args.predict_m, d))
                rec = {}
                for _ in range(2):   # first pass warms the workspaces
                    model = ctx.fit(data, spec)
                    fit_ph = ctx.timings()
                    t0 = time.perf_counter()
                    mu, cov, lo, hi = ctx.predict(model, Xs, want_cov=True, want_bars=True, sigma=2.0)
                    pred_wall = time.perf_counter() - t0
                    ph = ctx.timings()
                    ctx.model_destroy(model)
                m_ = args.predict_m
                dg_ = np.diag(cov)
                rec = {"m": m_, "fit_device_ms": fit_ph["total"], "predict_device_ms": ph["total"],
                       "predict_wall_ms_with_d2h": pred_wall * 1e3,
                       "cross_assemble_ms": ph["cross_assemble"], "trsm_ms": ph["trsm"], "syrk_ms": ph["syrk"],
                       "trsm_tflops": n * n * m_ / ph["trsm"] / 1e9, "trsm_frac": n * n * m_ / ph["trsm"] / 1e9 / dmma,
                       "syrk_tflops": m_ * m_ * n / ph["syrk"] / 1e9, "syrk_frac": m_ * m_ * n / ph["syrk"] / 1e9 / dmma,
                       "d2h_bytes": int(8 * (m_ * m_ + 3 * m_)),
                       "checks": {"finite": bool(np.isfinite(mu).all() and np.isfinite(cov).all() and np.isfinite(lo).all()),
                                  "variances_positive": bool((dg_ > 0).all()),
                                  "covariance_symmetric": bool(np.array_equal(cov, cov.T)),
                                  "bars_bracket_mean": bool((lo <= mu).all() and (mu <= hi).all()),
                                  # the reference's bars are mean -/+ sigma * row sums of chol(Sigma*), not 2 sigma_i
                                  "bars_symmetric_about_mean": bool(np.allclose(hi - mu, mu - lo, rtol=1e-9, atol=1e-12))}}
                roofline["predict"] = rec
            except Exception as ex:
                roofline["predict"] = {"error": f"{type(ex).__name__}: {ex}"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_record([cpu_sample(args.workload, args.cpu_sample_n, args.cpu_sample_n3)])
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": "timed at N=1 only (rank 0 of the single-GPU run)"}
    ctx.data_destroy(data)
    ctx.trim()
    barrier()

    # ---- the sharded north-star configurations, at this N ----------------------------------------------
    dmma_peak = D.bcast(peaks["dmma_tflops"] if peaks else None)
    legs = {}
    if args.c4_n > 0:
        try:
            legs["dist_c4"] = run_c4(ctx, D, rank, args, dmma_peak)
        except Exception as ex:  # reported, never hidden: the headline line must still come out
            legs["dist_c4"] = {"error": f"{type(ex).__name__}: {ex}"}
        ctx.trim()
    if args.c5_n > 0:
        try:
            legs["c5"] = run_c5(ctx, D, rank, args)
        except Exception as ex:
            legs["c5"] = {"error": f"{type(ex).__name__}: {ex}"}
        ctx.trim()
    barrier()
    D.close()
    if line is not None:
        line["roofline"].update(legs)   # last keys of the last object: they end up in the tail of stdout
        rl = line.pop("roofline")
        line["roofline"] = rl
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
