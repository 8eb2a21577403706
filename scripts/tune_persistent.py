"""Persistent (cross-tile pipelined) GEMM vs one tile per CTA: trailing-update SYRK at several K, a long-K NN
product, small-K panel shapes, and the whole C3 step."""
import sys, math
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
rng = np.random.default_rng(0)
n = 12288
C0 = np.zeros((n, n))
A768 = rng.standard_normal((n, 768))
A2 = rng.standard_normal((4096, 4096)); B2 = rng.standard_normal((4096, 4096))
P = rng.standard_normal((16384, 128)); W = rng.standard_normal((128, 128))
ref = {}
for pers in (0, 1, 0, 1):
    ctx.set_option("gemm_persistent", pers)
    out = []
    for k in (128, 256, 768):
        A = A768[:, :k]
        ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
        ms = min(ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)[1] for _ in range(3))
        out.append(f"syrk K={k}: {ms:.3f} ms {n*(n+128)*k/ms/1e9:.2f} TF/s")
    C2, _ = ctx.test_gemm(A2, B2)
    ms2 = min(ctx.test_gemm(A2, B2)[1] for _ in range(3))
    Cp, _ = ctx.test_gemm(P, W, transB=True)
    msp = min(ctx.test_gemm(P, W, transB=True)[1] for _ in range(3))
    key = "nn"
    if key not in ref: ref[key] = (C2, Cp)
    d = max(np.abs(C2 - ref[key][0]).max(), np.abs(Cp - ref[key][1]).max())
    print(f"persistent={pers}: " + " | ".join(out) + f" | nn 4096 {ms2:.3f} ms {2*4096**3/ms2/1e9:.2f} TF/s | panel 16384x128x128 {msp*1e3:.1f} us | maxdiff {d:.1e}", flush=True)
from oracle import gp_oracle as orc
X, y, _ = orc.synthetic(32768, 16, 0)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)
for pers in (0, 1, 0, 1):
    ctx.set_option("gemm_persistent", pers)
    ctx.energy_grad(data, spec)
    e, g = ctx.energy_grad(data, spec)
    tm = ctx.timings()
    print(f"persistent={pers}: C3 energy+grad total {tm['total']:.1f} ms potrf {tm['potrf']:.1f} inverse {tm['inverse']:.1f} solves {tm['solves']:.2f}  E={e!r}", flush=True)
