"""Assembly kernels side by side: exact differences (1), lean Gram (2), checked Gram (3).
Lower-triangle bytes / time against the measured HBM write peak."""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import dynaml_b200 as dg

ctx = dg.default_context()
peaks = ctx.measure_peaks() if hasattr(ctx, "measure_peaks") else None
out = {}
cases = [("matern52_d16", 16, lambda: dg.GenericMaternKernel(4.0, 2)._spec(noise=0.1)),
         ("se_d8", 8, lambda: dg.SEKernel(2.0, 1.0)._spec(noise=0.1)),
         ("ard_d8", 8, lambda: dg.MahalanobisKernel(np.full(8, 2.0), 1.0)._spec(noise=0.1)),
         ("se_d16", 16, lambda: dg.SEKernel(3.0, 1.0)._spec(noise=0.1)),
         ("matern32_d8", 8, lambda: dg.GenericMaternKernel(3.0, 1)._spec(noise=0.1))]
for n in (16384, 32768):
    for name, d, mk in cases:
        rng = np.random.default_rng(0)
        X = rng.standard_normal((n, d))
        y = rng.standard_normal(n)
        data = ctx.data_create(X, y)
        spec = mk()
        row = {}
        for mode in (1, 2, 3):
            ctx.set_option("assembly", mode)
            ctx.time_phase(data, spec, 0, 2)
            ms = ctx.time_phase(data, spec, 0, 10)
            row[mode] = ms
        ctx.set_option("assembly", 0)
        gb = 8 * n * (n + 1) / 2 / 1e9
        out[f"{name}_n{n}"] = {"ms_direct": row[1], "ms_lean": row[2], "ms_checked": row[3],
                               "lean_GBps": gb / row[2] * 1e3}
        print(f"N={n} {name}: direct {row[1]:.3f} ms  lean {row[2]:.3f} ms ({gb / row[2] * 1e3:.0f} GB/s)  "
              f"checked {row[3]:.3f} ms", flush=True)
        ctx.data_destroy(data)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/assemble_v3.json").write_text(json.dumps(out, indent=1))
