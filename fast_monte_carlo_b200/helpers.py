"""`sim_helpers.py` of the reference on the B200 engine: the temperature-softmax pass-outcome wrapper and the
quantile-interpolated yardage sampler (sim_helpers.py:4-38), with the tree ensembles evaluated by
`fmc_tree_predict` (CUDA) instead of xgboost / scikit-learn.

The reference classes take one-row DataFrames that go through the models' own preprocessors; these take the 17
numerics of FMC:676-682 as `[n, 17]` rows (the one-hot columns are the engine's `player`, "Unknown" by default).
The arithmetic after the margins is NumPy's, exactly as in the reference.  Inside a simulation the same sampler
runs in the kernel (`Engine(sampler="quantile_interp")`).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import Engine


def softmax(z: np.ndarray) -> np.ndarray:
    """Row-wise softmax with the max subtracted (sim_helpers.py:4-7)."""
    z = z - z.max(axis=1, keepdims=True)
    ez = np.exp(z)
    return ez / ez.sum(axis=1, keepdims=True)


class PassOutcomeModel:
    """sim_helpers.PassOutcomeModel (sim_helpers.py:9-24): softmax(margins / T) over the boosting rounds
    [0, best_iteration + 1) of a multi-class booster loaded in the engine."""

    def __init__(self, engine: Engine, model: str = "pass_stage2", temperature: float = 1.0,
                 best_iteration: Optional[int] = None):
        self.engine, self.model = engine, model
        self.T = float(temperature)
        f = engine.models[model]
        self.best_it = best_iteration if best_iteration is not None else f.best_iteration
        self.n_classes = f.n_outputs

    def predict_proba(self, rows17: np.ndarray) -> np.ndarray:
        tree_end = -1 if self.best_it is None else ((self.best_it or 0) + 1) * self.n_classes   # iteration_range
        margin = self.engine.predict(self.model, np.asarray(rows17, dtype=np.float64), 0, tree_end).astype(np.float32)
        return softmax(margin / self.T)


class QuantileYards:
    """sim_helpers.QuantileYards (sim_helpers.py:26-38): piecewise-linear inverse CDF through (q10, q50, q90),
    plus Normal(0, noise), clipped to [lo, hi]."""

    def __init__(self, engine: Engine, family: str):
        if family not in ("pass_yards", "run_yards", "sack_yards"):
            raise ValueError("family must be pass_yards, run_yards or sack_yards")
        self.engine, self.family = engine, family

    def quantiles(self, rows17: np.ndarray) -> np.ndarray:
        """[n, 3] = q10, q50, q90 (Pipeline.predict of the three GradientBoostingRegressors)."""
        return self.engine.predict(self.family, np.asarray(rows17, dtype=np.float64))

    def sample(self, row17: np.ndarray, lo: float, hi: float, noise: float = 0.5, rng=None) -> float:
        q10, q50, q90 = (float(v) for v in self.quantiles(np.asarray(row17, dtype=np.float64).reshape(1, -1))[0])
        r = np.random if rng is None else rng
        u = r.rand() if rng is None else rng.random()
        y = q10 + (q50 - q10) * (u / 0.5) if u < 0.5 else q50 + (q90 - q50) * ((u - 0.5) / 0.5)
        z = r.normal(0, noise)
        return float(np.clip(y + z, lo, hi))
