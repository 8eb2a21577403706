"""Host-side helpers either side of the hot path, mirroring the pieces of the
reference's housing workflow that feed and score the GP model:

  gaussianScalingTrainTest   CORE/DynaMLPipe.scala:784 (features and targets standardised with the
                             training mean / sample standard deviation)
  RegressionMetrics          CORE/evaluation/RegressionMetrics.scala (mae, rmse, Rsq, corr, yi)
  gpTuning                   CORE/DynaMLPipe.scala:1032-1083 ("GS" | "ML" | "CSA" switch)

These are O(N d) host computations; the model itself runs on the GPU through dynaml_b200.models."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .optimization import CoupledSimulatedAnnealing, GradBasedGlobalOptimizer, GridSearch


def gaussianScalingTrainTest(train: Sequence[Tuple[Sequence[float], float]], test: Sequence[Tuple[Sequence[float], float]]):
    """Returns (scaled train, scaled test, (feature mean, feature std, target mean, target std))."""
    Xtr = np.asarray([t[0] for t in train], dtype=np.float64)
    ytr = np.asarray([t[1] for t in train], dtype=np.float64)
    Xte = np.asarray([t[0] for t in test], dtype=np.float64)
    yte = np.asarray([t[1] for t in test], dtype=np.float64)
    mx, sx = Xtr.mean(0), Xtr.std(0, ddof=1)
    my, sy = float(ytr.mean()), float(ytr.std(ddof=1))
    tr = list(zip((Xtr - mx) / sx, (ytr - my) / sy))
    te = list(zip((Xte - mx) / sx, (yte - my) / sy))
    return tr, te, (mx, sx, my, sy)


class RegressionMetrics:
    """scoresAndLabels: (prediction, actual) pairs."""

    def __init__(self, scoresAndLabels: Sequence[Tuple[float, float]], length: int | None = None):
        p = np.asarray([s[0] for s in scoresAndLabels], dtype=np.float64)
        a = np.asarray([s[1] for s in scoresAndLabels], dtype=np.float64)
        self.length = len(p) if length is None else length
        err = p - a
        self.mae = float(np.abs(err).sum() / self.length)
        self.rmse = float(math.sqrt((err * err).sum() / self.length))
        self.rmsle = float(math.sqrt(np.sum((np.log(1 + np.abs(p)) - np.log(1 + np.abs(a))) ** 2) / self.length))
        ss_tot = float(((a - a.mean()) ** 2).sum())
        self.Rsq = float(1.0 - (err * err).sum() / ss_tot) if ss_tot > 0 else float("nan")
        sp, sa = p.std(), a.std()
        self.corr = float(((p - p.mean()) * (a - a.mean())).mean() / (sp * sa)) if sp > 0 and sa > 0 else float("nan")
        self.predictionEfficiency = self.Rsq
        self.sigma = float(math.sqrt(ss_tot / (self.length - 1.0))) if self.length > 1 else float("nan")

    def kpi(self) -> List[float]:
        return [self.mae, self.rmse, self.Rsq, self.corr]


def gpTuning(startingState: Dict[str, float], globalOpt: str = "GS", grid: int = 3, step: float = 0.02,
             maxIt: int = 20, policy: str = "GS", prior=None, rng=None):
    """DynaMLPipe.gpTuning: model => (tuned model, best configuration)."""
    def run(model):
        if globalOpt == "GS":
            opt = GridSearch(model).setGridSize(grid).setStepSize(step).setLogScale(False)
        elif globalOpt == "CSA":
            opt = CoupledSimulatedAnnealing(model, rng=rng).setGridSize(grid).setStepSize(step).setLogScale(False)
            opt.setMaxIterations(maxIt)
        elif globalOpt == "ML":
            return GradBasedGlobalOptimizer(model).optimize(
                startingState, {"tolerance": "0.0001", "step": str(step), "maxIterations": str(maxIt)})
        else:
            raise ValueError(f"unknown optimiser {globalOpt!r} (GS | CSA | ML)")
        if prior:
            opt.setPrior(prior)
        return opt.optimize(startingState, {"persist": "true"})
    return run
