"""The two GEMM kernels (option gemm_variant: 2 = cp.async kernel, 4 = bulk-copy / mbarrier kernel on the layouts of
gemm_v4_mask): correctness of all four operand layouts against NumPy, then trailing-update SYRK at several K, long-K
NN / TN products, a panel shape and the whole C3 step.  usage: tune_variant.py [2,4]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

import os
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 4]
ctx = dg.default_context()
for key in ("gemm_v4_mask", "gemm_v4_bk32_mask"):  # e.g. gemm_v4_mask=15 python scripts/tune_variant.py
    if key in os.environ:
        ctx.set_option(key, int(os.environ[key]))
rng = np.random.default_rng(0)

# ---- correctness: every layout, beta != 0, lower_only, small and ragged-in-tiles sizes -------------------
for v in variants:
    ctx.set_option("gemm_variant", v)
    worst = 0.0
    for (M, N, K) in ((128, 64, 32), (256, 192, 96), (384, 384, 160), (640, 256, 1024)):
        for tA in (False, True):
            for tB in (False, True):
                A = rng.standard_normal((K, M) if tA else (M, K))
                B = rng.standard_normal((N, K) if tB else (K, N))
                C0 = rng.standard_normal((M, N))
                opA = A.T if tA else A
                opB = B.T if tB else B
                want = 0.5 * (opA @ opB) - 1.5 * C0
                got, _ = ctx.test_gemm(A, B, C0, alpha=0.5, beta=-1.5, transA=tA, transB=tB)
                worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
                got0, _ = ctx.test_gemm(A, B, transA=tA, transB=tB)
                worst = max(worst, np.abs(got0 - opA @ opB).max() / np.abs(want).max())
    A = rng.standard_normal((512, 96)); C0 = rng.standard_normal((512, 512))
    got, _ = ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    want = C0 - A @ A.T
    il = np.tril_indices(512)
    worst = max(worst, np.abs(got[il] - want[il]).max())
    print(f"variant {v}: max rel err over layouts {worst:.2e}", flush=True)

# ---- timing -------------------------------------------------------------------------------------------
n = 12288
C0 = np.zeros((n, n))
A768 = rng.standard_normal((n, 768))
A2 = rng.standard_normal((4096, 4096)); B2 = rng.standard_normal((4096, 4096))
P = rng.standard_normal((16384, 128)); W = rng.standard_normal((128, 128))
for v in variants + variants:
    ctx.set_option("gemm_variant", v)
    out = []
    for k in (128, 256, 768):
        A = A768[:, :k]
        ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
        ms = min(ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)[1] for _ in range(3))
        out.append(f"syrk K={k}: {n*(n+128)*k/ms/1e9:.2f}")
    ctx.test_gemm(A2, B2)
    ms2 = min(ctx.test_gemm(A2, B2)[1] for _ in range(3))
    ctx.test_gemm(A2, B2, transA=True)
    ms3 = min(ctx.test_gemm(A2, B2, transA=True)[1] for _ in range(3))
    msp = min(ctx.test_gemm(P, W, transB=True)[1] for _ in range(3))
    print(f"variant {v}: " + " | ".join(out) + f" | nn 4096 {2*4096**3/ms2/1e9:.2f} | tn 4096 {2*4096**3/ms3/1e9:.2f} TF/s"
          f" | panel 16384x128x128 {msp*1e3:.1f} us", flush=True)

from oracle import gp_oracle as orc
X, y, _ = orc.synthetic(32768, 16, 0)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)
for v in variants + variants:
    ctx.set_option("gemm_variant", v)
    ctx.energy_grad(data, spec)
    e, g = ctx.energy_grad(data, spec)
    tm = ctx.timings()
    print(f"variant {v}: C3 energy+grad total {tm['total']:.1f} ms potrf {tm['potrf']:.1f} inverse {tm['inverse']:.1f}"
          f" E={e!r} g={list(g)!r}", flush=True)
