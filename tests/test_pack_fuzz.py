"""Random forests through the C++ specialiser/packer: the packed tables, walked in NumPy exactly like the
kernels walk them, must reproduce the oracle's margins on the ORIGINAL trees bit for bit -- for both tree
kinds, with and without "zero is missing", in both presets, including the layout corner cases (trees that
fold to constants, runs of more than 255 constants, all-constant outputs, empty tree ranges, depth 15)."""
import numpy as np
import pytest

import packed_walk as pw
from fast_monte_carlo_b200 import artifacts as art, native
from oracle import tree_oracle as to
from test_pack import _rows


def random_forest(rng, kind, n_trees, n_outputs, max_depth, zero_is_missing, n_onehot=6, p_leaf=0.25,
                  const_fraction=0.0):
    """A Forest over [n_onehot one-hot columns | 17 numerics] with random splits on realistic thresholds."""
    feat, thr, left, right, dl, val, roots, outs = [], [], [], [], [], [], [], []
    num_base = n_onehot

    def grow(depth, force_leaf):
        i = len(feat)
        feat.append(-1); thr.append(0.0); left.append(-1); right.append(-1); dl.append(0); val.append(0.0)
        if force_leaf or depth >= max_depth or (depth > 0 and rng.random() < p_leaf):
            v = rng.normal(0, 1.5)
            val[i] = float(np.float32(v)) if kind == art.KIND_XGB else float(v)
            return i
        if rng.random() < 0.15:
            c = int(rng.integers(0, n_onehot))
            feat[i] = c
            thr[i] = 0.5 if kind == art.KIND_SKL else float(np.float32(2.00001 if zero_is_missing else 0.5))
        else:
            k = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16]))
            feat[i] = num_base + k
            t = {0: rng.choice([1.5, 2.5, 3.5, 4.5]), 1: np.round(rng.uniform(0, 20), 1) + 0.05, 2: rng.integers(0, 100) + 0.5,
                 4: rng.integers(-21, 22) + (0.0 if rng.random() < 0.3 else 0.5), 5: rng.integers(0, 3600) + 0.5,
                 6: rng.choice([0.5, 2.5, 3.5]), 15: 1.5}.get(k)
            if t is None:
                t = rng.choice([0.5, 0.0, 1.0]) if k in (3, 12, 13, 14, 16) else np.round(rng.normal(5, 15), 1)
            thr[i] = float(np.float32(t))
        dl[i] = int(rng.integers(0, 2))
        l = grow(depth + 1, False)
        r = grow(depth + 1, False)
        left[i], right[i] = l, r
        return i

    for t in range(n_trees):
        roots.append(len(feat))
        outs.append(t % n_outputs)
        grow(0, rng.random() < const_fraction)
    cats = [f"c{j}" for j in range(n_onehot - 1)] + ["Unknown"]
    f = art.Forest(
        name="fuzz", kind=kind, link=art.LINK_SOFTMAX if n_outputs > 1 else art.LINK_SIGMOID, n_outputs=n_outputs,
        n_features=num_base + 17, num_base=num_base, n_num=17, zero_is_missing=bool(zero_is_missing),
        base_margin=rng.normal(0, 1, n_outputs).astype(np.float32).astype(np.float64),
        scale=0.1 if kind == art.KIND_SKL else 1.0, groups=[art.OneHotGroup("player", 0, cats)],
        feat=np.asarray(feat, np.int32), thr=np.asarray(thr, np.float32), left=np.asarray(left, np.int32),
        right=np.asarray(right, np.int32), default_left=np.asarray(dl, np.uint8), value=np.asarray(val, np.float64),
        tree_root=np.asarray(roots, np.int32), tree_out=np.asarray(outs, np.int32))
    return f      # (depth-first node order: not XGBoost's own layout, which the packer does not rely on)


def _check(f, mode, rows, cols, fold=None, tb=0, te=-1):
    skl = f.kind == art.KIND_SKL
    zm = bool(f.zero_is_missing) and not skl
    slots, stream, consts, meta = native.pack_forest_host(f, mode=mode, cols=cols, fold_values=fold, tree_begin=tb, tree_end=te)
    feats = pw.sim_rows(rows, zm) if mode == 0 else pw.predict_rows(rows, zm, xgb=not skl)
    got = pw.walk(slots, stream, consts, meta, feats, skl, f.base_margin)
    act = np.tile(np.array([cols[0]]), (rows.shape[0], 1))
    ref = to.raw_margin(f, rows, act, tb, f.n_trees if te < 0 else te)
    assert np.array_equal(got, ref)
    return meta


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("kind,zm", [(art.KIND_XGB, True), (art.KIND_XGB, False), (art.KIND_SKL, False)])
def test_random_forests_both_presets(native_lib, seed, kind, zm):
    rng = np.random.default_rng(1000 * kind + 10 * seed + int(zm))
    f = random_forest(rng, kind, n_trees=int(rng.integers(5, 60)), n_outputs=int(rng.integers(1, 4)),
                      max_depth=int(rng.integers(1, 8)), zero_is_missing=zm, const_fraction=float(rng.choice([0, 0.3])))
    rows = _rows(48, seed)
    hot = int(rng.integers(-1, 6))
    _check(f, 1, rows, (hot, -1))
    fold = np.zeros(17)
    fold[6] = fold[7] = 3.0
    fold[8:12] = [15.6, 35.7, 20.6, 0.0]          # an SP+ rating of exactly 0 (UTSA) folds as "missing" for CSR-fed boosters
    rows2 = rows.copy()
    rows2[:, 6:12] = fold[6:12]
    _check(f, 0, rows2, (hot, -1), fold)
    # iteration_range-style tree ranges, including the empty one
    _check(f, 1, rows, (hot, -1), tb=1, te=max(1, f.n_trees // 2))
    _check(f, 1, rows[:4], (hot, -1), tb=2, te=2)


def test_long_constant_runs_and_all_constant_outputs(native_lib):
    """> 255 consecutive constant trees need padding trees to carry them; an output whose trees all fold
    away is a pure sum of constants."""
    rng = np.random.default_rng(5)
    for kind in (art.KIND_XGB, art.KIND_SKL):
        f = random_forest(rng, kind, n_trees=900, n_outputs=3, max_depth=3, zero_is_missing=False, const_fraction=0.97)
        meta = _check(f, 1, _rows(16, 3), (-1, -1))
        assert meta["constants"] > 800
        g = random_forest(rng, kind, n_trees=40, n_outputs=2, max_depth=2, zero_is_missing=False, const_fraction=1.0)
        meta = _check(g, 1, _rows(16, 4), (-1, -1))
        assert meta["constants"] == 40 and meta["n_groups"][0] >= 1


def test_depth_limit(native_lib):
    rng = np.random.default_rng(6)
    ok = random_forest(rng, art.KIND_XGB, n_trees=3, n_outputs=1, max_depth=15, zero_is_missing=True, p_leaf=0.45)
    meta = _check(ok, 1, _rows(32, 5), (-1, -1))
    assert meta["max_depth"] <= 15
    # a 16-level chain is refused with a capacity error, not mis-evaluated
    n = 17
    feat = [6 + 2] * (n - 1) + [-1] + [-1] * (n - 1)
    f = art.Forest(
        name="deep", kind=art.KIND_XGB, link=art.LINK_SIGMOID, n_outputs=1, n_features=23, num_base=6, n_num=17,
        zero_is_missing=False, base_margin=np.zeros(1), scale=1.0, groups=[art.OneHotGroup("player", 0, list("abcdeU"))],
        feat=np.asarray(feat, np.int32), thr=np.asarray([float(i) for i in range(n - 1)] + [0.0] * n, np.float32),
        left=np.asarray([i + 1 for i in range(n - 1)] + [-1] * n, np.int32),
        right=np.asarray([n + i for i in range(n - 1)] + [-1] * n, np.int32),
        default_left=np.zeros(2 * n - 1, np.uint8), value=np.asarray([0.0] * (n - 1) + [float(i + 1) for i in range(n)], np.float64),
        tree_root=np.asarray([0], np.int32), tree_out=np.asarray([0], np.int32))
    with pytest.raises(native.FmcError, match="deeper"):
        native.pack_forest_host(f, mode=1, cols=(-1, -1))
