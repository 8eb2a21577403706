"""Device and float64 oracle against the extended-precision arbiter (oracle/extended.py) at the stress noise.
usage: arbitrate.py [n]"""
import sys, math, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dynaml_b200 as dg
from oracle import gp_oracle as orc, extended as ext

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
d = 8
X, y, _ = orc.synthetic(n, d, seed=n)
cases = {
    "se": (dg.SEKernel(1.5 * math.sqrt(d), 1.0), orc.Kernel("se", {"bandwidth": 1.5 * math.sqrt(d), "amplitude": 1.0})),
    "matern52": (dg.GenericMaternKernel(math.sqrt(d), 2), orc.Kernel("matern", {"p": 2.0, "l": math.sqrt(d)})),
}
for name, (k, ok) in cases.items():
    for noise in (0.1, 1e-3):
        model = dg.GPRegression(k, dg.DiracKernel(noise), list(zip(X, y)))
        e = model.energy(model._current_state)
        g = model.gradEnergy(model._current_state)
        ee, ge = ext.energy_grad(ok, noise, X, y)
        eo, go = orc.energy(ok, noise, X, y), orc.grad_energy(ok, noise, X, y)
        rec = {"kernel": name, "n": n, "noise": noise,
               "E_dev_rel": float(abs(e - ee) / abs(ee)), "E_orc_rel": float(abs(eo - ee) / abs(ee))}
        for key in go:
            rec[f"g[{key}]"] = {"value": float(ge[key]), "dev_rel": float(abs(g[key] - ge[key]) / abs(ge[key])),
                               "orc_rel": float(abs(go[key] - ge[key]) / abs(ge[key]))}
        print(json.dumps(rec), flush=True)
