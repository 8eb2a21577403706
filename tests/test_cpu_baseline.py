"""The timed CPU arm (oracle/cpu_baseline.py, bench.py --impl reference and `cpu_baseline`) computes what the
oracle computes: same energy and gradient as gp_oracle.energy / grad_energy for the three bench workloads."""
import math

import numpy as np
import pytest

from oracle import cpu_baseline as cb
from oracle import gp_oracle as orc


@pytest.mark.parametrize("kind,d", [("matern", 16), ("ard", 8), ("se", 8)])
def test_port_matches_oracle(kind, d):
    X, y, _ = orc.synthetic(300, d, seed=3)
    cov = cb.make_kernel(kind, d)
    tm = {}
    e, g = cb.energy_grad_fast(cov, 0.1, X, y, tm)
    eo = orc.energy(cov, 0.1, X, y)
    go = orc.grad_energy(cov, 0.1, X, y)
    assert e == pytest.approx(eo, rel=1e-10)
    assert set(g) == set(go)
    for k, v in go.items():
        assert g[k] == pytest.approx(v, rel=1e-8, abs=1e-8 * max(abs(x) for x in go.values()))
    assert tm["n2"] > 0 and tm["n3"] > 0


def test_port_reports_non_pd_like_the_reference():
    X, y, _ = orc.synthetic(50, 2, seed=1)
    cov = orc.Kernel("se", {"bandwidth": float("nan"), "amplitude": 1.0})
    e, g = cb.energy_grad_fast(cov, 0.1, X, y)
    assert e == float("inf") and all(math.isnan(v) for v in g.values())


@pytest.mark.parametrize("kind,d", [("matern", 16), ("se", 8)])
def test_timing_matrix_is_the_workload_matrix(kind, d):
    """The Gram-form matrix the LAPACK part is timed on equals the direct-difference one to rounding."""
    n = 200
    K, y = cb.timing_matrix(n, d, kind, seed=0)
    X, y2, _ = orc.synthetic(n, d, seed=0)
    Ko = orc.build_kernel_matrix(cb.make_kernel(kind, d), X, orc.dirac(0.1))
    assert np.array_equal(y, y2)
    assert np.abs(K - Ko).max() <= 1e-9


def test_time_parts_and_extrapolation():
    s = cb.time_parts(128, 256, 4, "se")
    assert s["n2"] == 128 and s["n3"] == 256 and s["t2"] > 0 and s["t3"] > 0 and s["threads"] >= 1
    assert cb.extrapolate(2.0, 3.0, 100, 200) == pytest.approx(2.0 * 4 + 3.0 * 8)
    assert cb.extrapolate(2.0, 3.0, 100, 400, n3_sample=200) == pytest.approx(2.0 * 16 + 3.0 * 8)
    assert cb.use_all_cores() >= 1


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs beside ours) on tiny samples: one JSON line with the same
    metric / unit / config as the GPU arm, the impl tag, a cpu_baseline record and an e2e object without copies."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-n", "256", "--cpu-sample-n3", "512"], capture_output=True, text=True, timeout=600,
                         cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"] == "GP energy+grad evals/s (N=32k FP64)" and line["unit"] == "evals/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["dtype"] == "f64"
    assert line["config"]["N"] == 32768 and line["config"]["d"] == 16
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["extrapolated"] is True and cpu["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
