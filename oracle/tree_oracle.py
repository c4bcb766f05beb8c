"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the tree arithmetic the reference delegates to third-party wheels:

  * XGBoost 3.0.4 CPU predictor behind `Booster.inplace_predict` / `Booster.predict`
    (call sites fast_monte_carlo_cfb.py:420, 745, 757; pass_outcome_infer.py:56-62).  xgboost is
    NOT installed in this image and is absent from /root/reference, so this part restates the
    published algorithm  ->  **parity unpinned** for the XGBoost boosters;
    the only anchors are the provisional vectors of SURVEY Appendix G (tests/golden/xgb_provisional.json).
  * scikit-learn 1.5.2 `GradientBoostingRegressor.predict` (`_predict_stages`) behind
    `Pipeline.predict` (fast_monte_carlo_cfb.py:784-809).  Pinned bit-for-bit against the live
    unpickled pipelines (tests/golden/make_golden.py -> tests/golden/sklearn_quantiles.npz).

Rows are described the way the reference's ColumnTransformer sees them: a [n,17] float64 block of
numerics in NUM order plus, per one-hot group, the absolute column that is hot (-1: the name is
not a category, OneHotEncoder(handle_unknown='ignore') emits all zeros).
"""
from __future__ import annotations

import numpy as np

KIND_XGB = 0
KIND_SKL = 1


def _feature_values(forest, cols: np.ndarray, num: np.ndarray, active: np.ndarray, rows: np.ndarray):
    """float32 value of column `cols[k]` for row `rows[k]` (+ a missing mask for CSR-fed boosters)."""
    nb = forest.num_base
    is_num = (cols >= nb) & (cols < nb + forest.n_num)
    v = np.zeros(cols.shape[0], dtype=np.float32)
    idx = np.clip(cols - nb, 0, forest.n_num - 1)
    v[is_num] = num[rows[is_num], idx[is_num]].astype(np.float32)
    if active.shape[1]:
        hot = np.zeros(cols.shape[0], dtype=bool)
        for g in range(active.shape[1]):
            hot |= (active[rows, g] == cols)
        v[~is_num] = hot[~is_num].astype(np.float32)
    return v


def leaf_values(forest, num: np.ndarray, active: np.ndarray, tree_begin: int = 0, tree_end: int | None = None,
                dense_nan_missing: bool = False) -> np.ndarray:
    """[n_rows, n_trees_in_range] float64 leaf value reached in every tree.

    XGBoost (Appendix D.1): leaf iff left == -1; missing -> default_left ? left : right; else
    f32(x) < f32(thr) ? left : right.  CSR-fed boosters (D.2): a value that is exactly 0 is absent
    from the CSR row and therefore *missing*.  sklearn (D.4): f32(x) <= thr ? left : right (thr was
    rounded toward -inf to f32 by the artifact compiler, which preserves the predicate exactly).
    """
    num = np.ascontiguousarray(num, dtype=np.float64)
    n = num.shape[0]
    active = np.asarray(active, dtype=np.int64).reshape(n, -1)
    tree_end = forest.n_trees if tree_end is None else tree_end
    roots = forest.tree_root[tree_begin:tree_end].astype(np.int64)
    nt = roots.shape[0]
    node = np.broadcast_to(roots, (n, nt)).copy().reshape(-1)
    rows = np.repeat(np.arange(n, dtype=np.int64), nt)
    live = np.flatnonzero(forest.left[node] >= 0)
    while live.size:
        nd = node[live]
        cols = forest.feat[nd].astype(np.int64)
        v = _feature_values(forest, cols, num, active, rows[live])
        thr = forest.thr[nd]
        if forest.kind == KIND_XGB:
            go_left = v < thr
            if forest.zero_is_missing:
                miss = v == 0.0
                go_left = np.where(miss, forest.default_left[nd] != 0, go_left)
            elif dense_nan_missing:
                miss = np.isnan(v)
                go_left = np.where(miss, forest.default_left[nd] != 0, go_left)
        else:
            go_left = v <= thr
        node[live] = np.where(go_left, forest.left[nd], forest.right[nd])
        live = live[forest.left[node[live]] >= 0]
    return forest.value[node].reshape(n, nt)


def raw_margin(forest, num, active, tree_begin: int = 0, tree_end: int | None = None) -> np.ndarray:
    """[n_rows, n_outputs] raw margins, accumulated in the model's own precision and tree order.

    XGBoost: float32, out = base_margin then += leaf for each tree in index order (tree t adds to
    output tree_out[t]).  sklearn: float64, out = init then += learning_rate * value per stage.
    """
    tree_end = forest.n_trees if tree_end is None else tree_end
    lv = leaf_values(forest, num, active, tree_begin, tree_end)
    n = lv.shape[0]
    outs = forest.tree_out[tree_begin:tree_end]
    res = np.zeros((n, forest.n_outputs), dtype=np.float64)
    for k in range(forest.n_outputs):
        sel = lv[:, outs == k]
        if forest.kind == KIND_XGB:
            seq = np.concatenate([np.full((n, 1), forest.base_margin[k], dtype=np.float32),
                                  sel.astype(np.float32)], axis=1)
            res[:, k] = np.cumsum(seq, axis=1, dtype=np.float32)[:, -1].astype(np.float64)
        else:
            seq = np.concatenate([np.full((n, 1), forest.base_margin[k], dtype=np.float64),
                                  forest.scale * sel], axis=1)
            res[:, k] = np.cumsum(seq, axis=1, dtype=np.float64)[:, -1]
    return res


def sigmoid_f32(m: np.ndarray) -> np.ndarray:
    """xgboost common::Sigmoid in float32: 1 / (expf(-x) + 1)."""
    m = np.asarray(m, dtype=np.float32)
    return (np.float32(1.0) / (np.exp(-m, dtype=np.float32) + np.float32(1.0))).astype(np.float32)


def softmax_f32(m: np.ndarray) -> np.ndarray:
    """xgboost common::Softmax: float32 expf(x - max), sum kept in double, divide by float(sum)."""
    m = np.asarray(m, dtype=np.float32)
    w = np.exp(m - m.max(axis=1, keepdims=True), dtype=np.float32)
    s = w.astype(np.float64).cumsum(axis=1)[:, -1].astype(np.float32)
    return (w / s[:, None]).astype(np.float32)


def play_model_features(forest, raw12: np.ndarray) -> np.ndarray:
    """[n,12] raw policy numerics -> the float64 row play_model.xgb was trained on: StandardScaler
    over 11 columns, is_red_zone raw (train_play_model.py:50-60 commented recipe)."""
    x = np.array(raw12, dtype=np.float64, copy=True)
    cols = forest.scaler_cols
    x[:, cols] = (x[:, cols] - forest.scaler_mean[None, :]) / forest.scaler_scale[None, :]
    return x
