"""First-contact diagnostics on a B200: peaks, GEMM correctness/perf, potrf, energy path."""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from dynaml_b200 import _capi
from oracle import gp_oracle as orc
import scipy.linalg as sla

out = {}
ctx = dg.default_context()
out["peaks"] = ctx.measure_peaks()
print("PEAKS", out["peaks"], flush=True)

rng = np.random.default_rng(0)
# --- GEMM correctness, all layouts
for (ta, tb) in [(0, 0), (0, 1), (1, 0), (1, 1)]:
    M, N, K = 200, 300, 150
    A = rng.standard_normal((K, M) if ta else (M, K)); B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    C, ms = ctx.test_gemm(A, B, C0, alpha=0.7, beta=-0.3, transA=bool(ta), transB=bool(tb))
    ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 0.3 * C0
    print("gemm", ta, tb, "maxerr", np.abs(C - ref).max(), flush=True)
# --- GEMM perf
for n in (4096, 8192):
    A = rng.standard_normal((n, 256)); 
    C, ms = ctx.test_gemm(A, A, np.zeros((n, n)), alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    C, ms = ctx.test_gemm(A, A, np.zeros((n, n)), alpha=-1.0, beta=1.0, transB=True, lower_only=True)
    fl = n * (n + 128) * 256  # lower tiles incl diag
    print(f"syrk n={n} k=256: {ms:.3f} ms  {fl/ms/1e9:.2f} TF/s", flush=True)
    out[f"syrk_{n}"] = fl / ms / 1e9
for n in (4096,):
    A = rng.standard_normal((n, n)); B = rng.standard_normal((n, n))
    C, ms = ctx.test_gemm(A, B); C, ms = ctx.test_gemm(A, B)
    print(f"gemm nn n={n}: {ms:.3f} ms {2*n**3/ms/1e9:.2f} TF/s err {np.abs(C - A@B).max():.3e}", flush=True)
    out[f"gemm_{n}"] = 2 * n ** 3 / ms / 1e9
# --- potrf
for n in (128, 256, 300, 1024, 4096, 8192):
    G = rng.standard_normal((n, n + 8)); S = G @ G.T + n * 1e-3 * np.eye(n)
    for la in (0, 1):
        ctx.set_option("lookahead", la)
        L, Li, Ai, ld, ms = ctx.test_potrf(S, want_linv=(n <= 4096), want_ainv=(n <= 4096))
        Lr = sla.cholesky(S, lower=True)
        e1 = np.abs(L - Lr).max() / np.abs(Lr).max()
        msg = f"potrf n={n} la={la}: {ms:.3f} ms {n**3/3/ms/1e9:.2f} TF/s relerr {e1:.2e} logdet err {abs(ld - np.log(np.diag(Lr)).sum()):.2e}"
        if Li is not None:
            msg += f" inv err {np.abs(Li @ Lr - np.eye(n)).max():.2e} ainv err {np.abs(Ai @ S - np.eye(n)).max():.2e}"
        print(msg, flush=True)
        out[f"potrf_{n}_la{la}"] = n ** 3 / 3 / ms / 1e9
# --- energy + grad vs oracle
X, y, Xs = orc.synthetic(1000, 8, 1, 64)
cases = {
  "se": (dg.SEKernel(2.0, 1.3), orc.Kernel("se", {"bandwidth": 2.0, "amplitude": 1.3})),
  "matern": (dg.GenericMaternKernel(2.5, 2), orc.Kernel("matern", {"p": 2.0, "l": 2.5})),
  "ard": (dg.MahalanobisKernel(np.linspace(1.5, 3, 8), 1.1), orc.Kernel("ard", {"MahalanobisAmplitude": 1.1, **{f"MahalanobisBandwidth_{i}": b for i, b in enumerate(np.linspace(1.5, 3, 8))}})),
  
}
for name, (k, ok) in cases.items():
    m = dg.GPRegression(k, dg.DiracKernel(0.3), list(zip(X, y)))
    e = m.energy(m._current_state); eo = orc.energy(ok, 0.3, X, y)
    g = m.gradEnergy(m._current_state); go = orc.grad_energy(ok, 0.3, X, y)
    Kd = k.buildKernelMatrix(X, len(X)).getKernelMatrix(); Ko = orc.build_kernel_matrix(ok, X)
    print(name, "K relerr", (np.abs(Kd - Ko) / np.abs(Ko).clip(1e-300)).max(), "energy", e, eo, "rel", abs(e - eo) / abs(eo), flush=True)
    for kk in go:
        print("   grad", kk, g.get(kk), go[kk], "rel", abs(g.get(kk, np.nan) - go[kk]) / max(abs(go[kk]), 1e-300), flush=True)
    pd = m.predictiveDistribution(Xs); mu, cv = orc.predictive(ok, 0.3, X, y, Xs)
    print("   pred mean rel", np.abs(pd.mu - mu).max() / np.abs(mu).max(), "cov rel", np.abs(pd.covariance - cv).max() / np.abs(cv).max(), flush=True)
    try:
        lo, hi = pd.underlyingDist.confidenceInterval(2.0); lo2, hi2 = orc.error_bars(mu, cv, 2.0)
        print("   bars rel", np.abs(lo - lo2).max() / np.abs(lo2).max(), flush=True)
    except Exception as ex:
        print("   bars failed:", ex, flush=True)
# --- timing of the big configs
for (n, d, kern) in [(8192, 8, "ard"), (16384, 8, "se"), (32768, 16, "matern")]:
    X, y, _ = orc.synthetic(n, d, 0)
    if kern == "ard": k = dg.MahalanobisKernel(np.full(d, np.sqrt(d)), 1.0)
    elif kern == "se": k = dg.SEKernel(np.sqrt(d), 1.0)
    else: k = dg.GenericMaternKernel(np.sqrt(d), 2)
    data = ctx.data_create(X, y)
    spec = k._spec(noise=0.1)
    for la in (0, 1):
        ctx.set_option("lookahead", la)
        for rep in range(2):
            t0 = time.time(); e = ctx.energy(data, spec); t1 = time.time()
        tm = ctx.timings()
        print(f"energy n={n} {kern} la={la}: {1e3*(t1-t0):.1f} ms wall, E={e:.6f}, timings {tm}, potrf {n**3/3/tm['potrf']/1e9:.2f} TF/s, assemble {8*n*(n+1)/2/tm['assemble']/1e6:.1f} GB/s", flush=True)
        out[f"energy_{n}_la{la}"] = tm
    t0 = time.time(); e, g = ctx.energy_grad(data, spec); t1 = time.time()
    tm = ctx.timings()
    print(f"energy_grad n={n} {kern}: {1e3*(t1-t0):.1f} ms wall, timings {tm}, total {n**3/tm['total']/1e9:.2f} TF/s, inverse {2*n**3/3/tm['inverse']/1e9:.2f} TF/s", flush=True)
    out[f"energy_grad_{n}"] = tm
    ctx.data_destroy(data)
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(out, open("gpurun_out/gpu_first.json", "w"), indent=1, default=float)
