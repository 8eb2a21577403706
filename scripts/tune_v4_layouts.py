"""Which operand layouts should use the bulk-copy GEMM kernel, and with which slice width: option gemm_v4_mask
(bit 2*a_kmajor + b_kmajor: 1 = NT, 2 = NN, 4 = TT, 8 = TN) and gemm_v4_bk32_mask (same bits: 32-wide slices, 2 stages).
usage: tune_v4_layouts.py [v4_mask,bk32_mask ...]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

ctx = dg.default_context()
rng = np.random.default_rng(0)
n = 12288
C0 = np.zeros((n, n)); A768 = rng.standard_normal((n, 768))
A2 = rng.standard_normal((4096, 4096)); B2 = rng.standard_normal((4096, 4096))
X, y, _ = orc.synthetic(32768, 16, 0)
data = ctx.data_create(X, y)
spec = dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)
ref = None
combos = ((0, 0), (1, 0), (1, 1), (3, 0), (3, 2), (15, 8), (15, 10))
if len(sys.argv) > 1:
    combos = tuple(tuple(int(v) for v in c.split(",")) for c in sys.argv[1:])
for (mask, bk32) in combos:
    ctx.set_option("gemm_v4_mask", mask); ctx.set_option("gemm_v4_bk32_mask", bk32)
    ctx.test_gemm(A768, A768, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    ms = min(ctx.test_gemm(A768, A768, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)[1] for _ in range(2))
    Cnn, _ = ctx.test_gemm(A2, B2); ms2 = min(ctx.test_gemm(A2, B2)[1] for _ in range(2))
    Ctn, _ = ctx.test_gemm(A2, B2, transA=True); ms3 = min(ctx.test_gemm(A2, B2, transA=True)[1] for _ in range(2))
    if ref is None: ref = (Cnn, Ctn)
    d = max(np.abs(Cnn - ref[0]).max(), np.abs(Ctn - ref[1]).max())
    ctx.energy_grad(data, spec)
    e, g = ctx.energy_grad(data, spec)
    tm = ctx.timings()
    print(f"v4_mask {mask:2d} bk32_mask {bk32:2d}: syrk768 {n*(n+128)*768/ms/1e9:.2f} | nn {2*4096**3/ms2/1e9:.2f} | tn {2*4096**3/ms3/1e9:.2f} TF/s"
          f" | maxdiff {d:.1e} | C3 total {tm['total']:.1f} potrf {tm['potrf']:.1f} inverse {tm['inverse']:.1f} ms E={e!r} g1={float(g[1])!r}", flush=True)
