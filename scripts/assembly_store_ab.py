"""Lean SE assembly: 16-byte st.global from registers against staged chunks + bulk (TMA) stores (option assembly_store).
Same entries bit for bit; time of the lower-triangle build at N = 32768, d = 8 and d = 16."""
import sys, json, math
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
ctx = dg.default_context()
rng = np.random.default_rng(0)
Xs = rng.standard_normal((1024, 8))
ref = None
for mode in (0, 1):
    ctx.set_option("assembly_store", mode)
    K = ctx.kernel_matrix(dg.SEKernel(2.0, 1.3)._spec(noise=0.2), Xs, add_noise=True)
    ref = K if ref is None else ref
    print("mode", mode, "identical to mode 0:", bool(np.array_equal(K, ref)), "symmetric:", bool(np.array_equal(K, K.T)))
n = 32768
for d in (8, 16):
    data = ctx.data_create(rng.standard_normal((n, d)), rng.standard_normal(n))
    spec = dg.SEKernel(math.sqrt(d), 1.0)._spec(noise=0.1)
    out = {"n": n, "d": d}
    for mode in (0, 1, 0, 1):
        ctx.set_option("assembly_store", mode)
        ms = ctx.time_phase(data, spec, 0, reps=20)
        out.setdefault(f"store={mode}", []).append(round(ms, 4))
    out["gbs"] = {k: round(8 * n * (n + 1) / 2 / min(v) / 1e6, 1) for k, v in out.items() if k.startswith("store")}
    print(json.dumps(out), flush=True)
    ctx.data_destroy(data)
ctx.set_option("assembly_store", 0)
