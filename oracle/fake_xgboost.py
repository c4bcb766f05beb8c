"""ORACLE (test infrastructure only).  A stand-in for the `xgboost` package, which is not installed
in this image and not vendored under /root/reference, so that the *unmodified* reference module
fast_monte_carlo_cfb.py can be imported and run here to generate golden trajectories
(oracle/ref_harness.py, tests/golden/make_golden.py).

It restates the published XGBoost 3.0.4 CPU prediction algorithm for
the calls the reference makes:
    xgb.Booster().load_model(path)                    FMC:329, 641-642
    Booster.set_param(...)                            FMC:330, 645-646
    Booster.inplace_predict(csr)  -> transformed      FMC:745, 757
    Booster.predict(DMatrix, output_margin=True, iteration_range=...)   FMC:420, sim_helpers.py:23
Rows arrive as scipy CSR (ColumnTransformer output): entries that are not stored are *missing*.
**Parity unpinned**: no real xgboost is available to check this file against.
"""
from __future__ import annotations

import json
import os

import numpy as np

__version__ = "3.0.4-oracle-standin"

# Stand-in for the stage-2 booster that the reference snapshot does not ship
# (.MISSING_LARGE_BLOBS).  Raw class probabilities [incomplete, intercepted, sack] before the
# reference's own nudges (FMC:764-770).  Same constants in the C oracle and the CUDA engine.
STAGE2_STANDIN_PROBS = (0.78, 0.05, 0.17)


class DMatrix:
    def __init__(self, data, enable_categorical=False, **kw):
        self.data = data


class Booster:
    def __init__(self, *a, **k):
        self.learner = None
        self.path = None

    def set_param(self, *a, **k):
        return None

    def load_model(self, path):
        self.path = str(path)
        if not os.path.exists(path):
            # the missing stage-2 blob: fall back to the documented stand-in
            self.learner = None
            return
        with open(path, "r") as f:
            self.learner = json.load(f)["learner"]
        m = self.learner["gradient_booster"]["model"]
        self._trees = []
        self._cats = []          # per tree: {node: set of categories that go RIGHT} (common/categorical.h Decision)
        for t in m["trees"]:
            self._trees.append((
                np.asarray(t["left_children"], dtype=np.int64),
                np.asarray(t["right_children"], dtype=np.int64),
                np.asarray(t["split_indices"], dtype=np.int64),
                np.asarray(t["split_conditions"], dtype=np.float64).astype(np.float32),
                np.asarray(t["default_left"], dtype=np.int64),
            ))
            sets = {}
            cats = t.get("categories", [])
            for j, node in enumerate(t.get("categories_nodes", [])):
                a, n = t["categories_segments"][j], t["categories_sizes"][j]
                sets[int(node)] = set(int(c) for c in cats[a:a + n])
            self._cats.append(sets)
        self._tree_info = list(m["tree_info"])
        lmp = self.learner["learner_model_param"]
        self._n_class = max(1, int(lmp.get("num_class", "0")))
        self._objective = self.learner["objective"]["name"]
        bs = np.float32(float(lmp["base_score"]))
        if self._objective == "binary:logistic":
            self._base = np.float32(-np.log(np.float32(1.0) / bs - np.float32(1.0), dtype=np.float32))
        else:
            self._base = bs

    # -- one sparse row -> raw margins (float32, sequential tree order) -------------------------
    def _margin_row(self, cols, vals, tree_begin=0, tree_end=None):
        present = {int(c): np.float32(v) for c, v in zip(cols, vals)}
        out = np.full(self._n_class, self._base, dtype=np.float32)
        tree_end = len(self._trees) if tree_end is None else tree_end
        for t in range(tree_begin, tree_end):
            lc, rc, si, sc, dl = self._trees[t]
            cat_sets = self._cats[t]
            i = 0
            while lc[i] != -1:
                v = present.get(int(si[i]))
                if v is None:
                    i = lc[i] if dl[i] else rc[i]
                elif i in cat_sets:      # categorical split: a category in the node's set goes right
                    i = rc[i] if int(v) in cat_sets[i] else lc[i]
                else:
                    i = lc[i] if v < sc[i] else rc[i]
            k = self._tree_info[t]
            out[k] = np.float32(out[k] + sc[i])
        return out

    def _transform(self, m):
        m = np.asarray(m, dtype=np.float32)
        if self._objective == "binary:logistic":
            return (np.float32(1.0) / (np.exp(-m, dtype=np.float32) + np.float32(1.0)))[:, 0]
        if self._objective.startswith("multi:"):
            w = np.exp(m - m.max(axis=1, keepdims=True), dtype=np.float32)
            s = np.float32(np.cumsum(w.astype(np.float64), axis=1)[:, -1])
            return w / s[:, None]
        return m

    def _rows(self, X):
        import scipy.sparse as sp
        if sp.issparse(X):
            X = X.tocsr()
            for r in range(X.shape[0]):
                a, b = X.indptr[r], X.indptr[r + 1]
                yield X.indices[a:b], X.data[a:b]
        else:
            A = np.asarray(X, dtype=np.float64)
            for r in range(A.shape[0]):
                keep = ~np.isnan(A[r])
                yield np.flatnonzero(keep), A[r][keep]

    def inplace_predict(self, X, iteration_range=None, **kw):
        if self.learner is None:
            n = X.shape[0]
            return np.tile(np.asarray(STAGE2_STANDIN_PROBS, dtype=np.float32), (n, 1))
        tb, te = 0, None
        if iteration_range is not None and iteration_range != (0, 0):
            tb, te = iteration_range[0] * self._n_class, iteration_range[1] * self._n_class
        m = np.stack([self._margin_row(c, v, tb, te) for c, v in self._rows(X)])
        return self._transform(m)

    def predict(self, dmat, output_margin=False, iteration_range=None, **kw):
        X = dmat.data if isinstance(dmat, DMatrix) else dmat
        if hasattr(X, "dtypes") and hasattr(X, "columns"):
            # DataFrame (enable_categorical=True): a categorical column is presented by its category CODE
            cols = []
            for c in X.columns:
                col = X[c]
                cols.append(col.cat.codes.to_numpy(dtype=np.float64) if str(col.dtype) == "category"
                            else col.to_numpy(dtype=np.float64))
            X = np.stack(cols, axis=1)
        elif hasattr(X, "to_numpy"):
            X = X.to_numpy(dtype=np.float64)
        tb, te = 0, None
        if iteration_range is not None and iteration_range != (0, 0):
            tb, te = iteration_range[0] * self._n_class, iteration_range[1] * self._n_class
        m = np.stack([self._margin_row(c, v, tb, te) for c, v in self._rows(X)])
        if output_margin:
            return m if self._n_class > 1 else m[:, 0]
        return self._transform(m)
