"""Small GEMM workload for ncu: the trailing-update SYRK (K = 256) and a long-K NN GEMM."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg

ctx = dg.default_context()
rng = np.random.default_rng(0)
n = 12288
A = rng.standard_normal((n, 256))
C0 = np.zeros((n, n))
for _ in range(3):
    C, ms = ctx.test_gemm(A, A, C0, alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    print(f"syrk n={n} k=256: {ms:.3f} ms {n*(n+128)*256/ms/1e9:.2f} TF/s")
n = 4096
A = rng.standard_normal((n, n)); B = rng.standard_normal((n, n))
for _ in range(2):
    C, ms = ctx.test_gemm(A, B)
    print(f"gemm nn n={n}: {ms:.3f} ms {2*n**3/ms/1e9:.2f} TF/s")
