"""Config C3 prediction: fit at N = 32768 (Matern-5/2, d = 16) then predictive mean + covariance + error bars
for M = 4096 test points; phase timings from the library's CUDA events."""
import sys, math, time, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = 16
ctx = dg.default_context()
X, y, Xs = orc.synthetic(n, d, 0, m)
spec = dg.GenericMaternKernel(math.sqrt(d), 2)._spec(noise=0.1)
data = ctx.data_create(X, y)
for rep in range(2):
    t0 = time.perf_counter(); model = ctx.fit(data, spec); t_fit = time.perf_counter() - t0
    fit_t = ctx.timings()
    t0 = time.perf_counter(); mean, cov, lo, hi = ctx.predict(model, Xs, want_cov=True, want_bars=True, sigma=2.0); t_pred = time.perf_counter() - t0
    tm = ctx.timings()
    if rep == 0: ctx.model_destroy(model)
print(json.dumps({"n": n, "m": m, "fit_ms": t_fit * 1e3, "fit_phases": fit_t, "predict_wall_ms": t_pred * 1e3, "predict_phases": tm,
                  "cross_assemble_gbs": 8 * n * m / tm["cross_assemble"] / 1e6, "trsm_tflops": n * n * m / tm["trsm"] / 1e9,
                  "syrk_tflops": m * m * n / tm["syrk"] / 1e9, "cov_psd_min_diag": float(np.diag(cov).min()),
                  "bars_finite": bool(np.isfinite(lo).all())}))
# size-independent check: posterior variance at the training points themselves is below the noise variance-ish
mu_tr, cov_tr, _, _ = ctx.predict(model, X[:256], want_cov=True)
print("train-point check: max |mu - y| =", float(np.abs(mu_tr - y[:256]).max()), " max var =", float(np.diag(cov_tr).max()))
ctx.model_destroy(model); ctx.data_destroy(data)
