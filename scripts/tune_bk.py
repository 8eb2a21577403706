"""GEMM K-slice (16 vs 32) on the trailing-update SYRK, a long-K NN product and the whole C3 step."""
import sys, math, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

ctx = dg.default_context()
rng = np.random.default_rng(0)
n = 12288
A = rng.standard_normal((n, 256)); C0 = np.zeros((n, n))
A2 = rng.standard_normal((4096, 4096)); B2 = rng.standard_normal((4096, 4096))
ref = None
for bk in (16, 32, 16, 32):
    ctx.set_option("gemm_bk", bk)
    ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    ms = min(ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)[1] for _ in range(3))
    C2, _ = ctx.test_gemm(A2, B2)
    ms2 = min(ctx.test_gemm(A2, B2)[1] for _ in range(3))
    if ref is None: ref = C2
    print(f"bk={bk}: syrk n={n} k=256 {ms:.3f} ms {n*(n+128)*256/ms/1e9:.2f} TF/s | nn 4096 {ms2:.3f} ms {2*4096**3/ms2/1e9:.2f} TF/s | maxdiff {np.abs(C2-ref).max():.2e}", flush=True)
X, y, _ = orc.synthetic(32768, 16, 0)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)
for bk in (16, 32, 16, 32):
    ctx.set_option("gemm_bk", bk)
    ctx.energy_grad(data, spec)
    e, g = ctx.energy_grad(data, spec)
    print(f"bk={bk}: C3 energy+grad {ctx.timings()}  E={e!r}", flush=True)
