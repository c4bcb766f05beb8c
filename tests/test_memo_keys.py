"""Exact-memo keys (csrc/fmc_memo.hpp) checked on the CPU against the oracle: states that share a key must share
their margins bit for bit -- that is the whole correctness argument of the memo (a hit returns what the walk would
have computed).  The keys come from the library's own host entry point (fmc_memo_keys_host: the same
build_rank_spec / memo_key code the kernel runs); the margins from the C oracle walking the ORIGINAL trees."""
import numpy as np
import pytest

from conftest import ISU, KSU
from fast_monte_carlo_b200 import artifacts as art, native

FAMS = ("pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards", "play_model")


def _states(rng, n):
    """Simulation-like states plus values sitting on / next to typical thresholds and exact zeros."""
    down = rng.choice([1, 2, 3, 4, 5, 6], size=n, p=[.36, .30, .20, .10, .03, .01]).astype(float)
    dist = np.where(rng.random(n) < 0.5, np.round(rng.normal(8, 4, n), 0), rng.normal(8, 5, n))
    dist = np.clip(dist, 0.0, 40.0)
    ytg = np.where(rng.random(n) < 0.5, rng.integers(0, 101, n).astype(float), rng.random(n) * 100.0)
    half = rng.random(n) < 0.15                      # x.5 and float32-rounded neighbours of x.5: where trees split
    ytg = np.where(half, np.floor(ytg) + 0.5 + rng.choice([0.0, 1e-7, -1e-7, 3e-5], size=n), ytg)
    dist = np.where(rng.random(n) < 0.15, np.floor(dist) + 0.5 + rng.choice([0.0, 1e-7, -1e-7], size=n), dist)
    dist[rng.random(n) < 0.02] = 0.0
    ytg[rng.random(n) < 0.02] = 0.0
    sd = np.round(rng.normal(0, 14, n))
    sd[rng.random(n) < 0.2] = 0.0
    sd[rng.random(n) < 0.01] = rng.choice([-300.0, 300.0, -256.0, 255.0])
    sec = rng.integers(1, 3601, n).astype(float)
    sec[rng.random(n) < 0.1] = rng.choice([1, 120, 121, 1800, 1801, 1920, 1921, 3600], size=1)[0]
    base = np.stack([down, dist, ytg, sd, sec], axis=1)
    # ten jittered copies of every base state: most stay inside the base state's rank cell (repeated keys), some
    # step over a threshold (a key must then change whenever an output may)
    rep = np.repeat(base, 10, axis=0)
    m = rep.shape[0]
    keep = rng.random(m) < 0.3
    rep[:, 1] = np.where(keep | (rep[:, 1] == 0.0), rep[:, 1], np.clip(rep[:, 1] + rng.normal(0, 0.05, m), 0.01, 40.0))
    rep[:, 2] = np.where(keep | (rep[:, 2] == 0.0), rep[:, 2], np.clip(rep[:, 2] + rng.normal(0, 0.05, m), 0.01, 100.0))
    rep[:, 4] = np.clip(rep[:, 4] + np.where(keep, 0, rng.integers(-2, 3, m)), 1, 3600)
    rep[:, 3] = rep[:, 3] + np.where(rng.random(m) < 0.1, 1.0, 0.0)
    return rep


def _rows17(st, sp_off, sp_def):
    """The 17 numerics `_fill_row` builds from a state (FMC:996-1021)."""
    n = st.shape[0]
    x = np.zeros((n, 17))
    down, dist, ytg, sd, sec = st.T
    x[:, 0], x[:, 1], x[:, 2], x[:, 4], x[:, 5] = down, dist, ytg, sd, sec
    x[:, 3] = ytg <= 20.0
    x[:, 6] = x[:, 7] = 3.0
    x[:, 8], x[:, 9], x[:, 10], x[:, 11] = sp_off[0], sp_off[1], sp_def[2], sp_def[0]
    x[:, 12] = dist >= ytg - 0.5
    x[:, 13] = (down == 4) & (dist <= 2.0)
    x[:, 14] = ytg <= 33.0
    x[:, 15] = np.where(sec > 1800, 1.0, 2.0)
    x[:, 16] = (sec % 1800) <= 120
    return x


@pytest.mark.parametrize("name", FAMS)
@pytest.mark.parametrize("orient", [0, 1])
def test_equal_keys_mean_equal_margins(native_lib, oracle, models_s2, name, orient):
    f = models_s2[name]
    fam = art.MODEL_IDS[name]
    sp = (KSU, ISU) if orient == 0 else (ISU, KSU)
    fold = np.zeros(17)
    fold[6] = fold[7] = 3.0
    fold[8], fold[9], fold[10], fold[11] = sp[0][0], sp[0][1], sp[1][2], sp[1][0]
    cols = [-1, -1]
    if name != "play_model":
        for gi, g in enumerate(f.groups[:2]):
            cols[gi] = g.column_of("Unknown")
    rng = np.random.default_rng(11 + fam * 2 + orient)
    st = _states(rng, 6_000)
    keys, meta = native.memo_keys_host(f, fam, st, cols=cols, fold_values=fold)
    if name == "play_model" and keys is None:
        pytest.skip("play_model.xgb: " + meta["why"])      # its rank vector may not fit a key: then it is always walked
    assert keys is not None, meta
    assert (keys >> np.uint64(63)).all()                   # the valid bit: 0 is the empty slot
    rows = _rows17(st, sp[0], sp[1])
    active = np.tile(np.asarray(cols, dtype=np.int32), (rows.shape[0], 1))
    if f.scaler_cols is not None:        # fo_predict, like Booster.predict, takes the standardised row (scaler.pkl)
        for j, c in enumerate(f.scaler_cols):
            rows[:, c] = (rows[:, c] - f.scaler_mean[j]) / f.scaler_scale[j]
    out = oracle.predict(name, rows[:, :f.n_num], active, f.n_outputs)
    order = np.argsort(keys, kind="stable")
    k, o = keys[order], out[order].view(np.uint64)
    same = k[1:] == k[:-1]
    assert same.sum() > 1000, "the sample must contain repeated keys for the check to mean anything"
    assert np.array_equal(o[1:][same], o[:-1][same]), f"{name}: two states with one key got different margins"
    # and the key is not trivially unique per state, nor one key for everything
    n_keys = len(np.unique(keys))
    assert 50 < n_keys < len(keys)


def test_keys_separate_what_the_forest_separates(native_lib, oracle, models_s2):
    """Sanity in the other direction: on a grid of states the number of distinct keys is at least the number of
    distinct outputs (a key may be finer than necessary, never coarser)."""
    f = models_s2["pass_yards"]
    rng = np.random.default_rng(5)
    st = _states(rng, 3_000)
    fold = np.zeros(17)
    fold[6] = fold[7] = 3.0
    fold[8], fold[9], fold[10], fold[11] = KSU[0], KSU[1], ISU[2], ISU[0]
    cols = [g.column_of("Unknown") for g in f.groups[:2]]
    keys, _ = native.memo_keys_host(f, art.MODEL_IDS["pass_yards"], st, cols=cols, fold_values=fold)
    rows = _rows17(st, KSU, ISU)
    out = oracle.predict("pass_yards", rows, np.tile(np.asarray(cols, np.int32), (len(rows), 1)), 3)
    n_out = len(np.unique(out.view(np.uint64), axis=0))
    assert len(np.unique(keys)) >= n_out


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("kind,zm", [(art.KIND_XGB, True), (art.KIND_XGB, False), (art.KIND_SKL, False)])
def test_random_forests_equal_keys_equal_margins(native_lib, seed, kind, zm):
    """The same property on random forests (both tree kinds, with and without "zero is missing", thresholds on either
    side of the 0/1 flags, splits on the folded numerics and on one-hot columns): whatever the specialiser keeps,
    the key must determine the margins."""
    from oracle import tree_oracle as to
    from test_pack_fuzz import random_forest
    rng = np.random.default_rng(7000 + 100 * kind + 10 * seed + int(zm))
    f = random_forest(rng, kind, n_trees=int(rng.integers(10, 80)), n_outputs=int(rng.integers(1, 4)),
                      max_depth=int(rng.integers(2, 7)), zero_is_missing=zm, const_fraction=float(rng.choice([0, 0.2])))
    fold = np.zeros(17)
    fold[6] = fold[7] = 3.0
    fold[8:12] = [15.6, 35.7, 20.6, 0.0]          # an SP+ rating of exactly 0 folds as "missing" for CSR-fed boosters
    hot = int(rng.integers(-1, 6))
    st = _states(rng, 3000)
    keys, meta = native.memo_keys_host(f, 0 if kind == art.KIND_XGB else 2, st, cols=(hot, -1), fold_values=fold)
    assert keys is not None, meta
    rows = _rows17(st, (fold[8], fold[9], 0.0), (fold[11], 0.0, fold[10]))
    out = to.raw_margin(f, rows, np.tile(np.array([hot]), (rows.shape[0], 1)))
    order = np.argsort(keys, kind="stable")
    k, o = keys[order], np.ascontiguousarray(out[order]).view(np.uint64)
    same = k[1:] == k[:-1]
    assert same.sum() > 500
    assert np.array_equal(o[1:][same], o[:-1][same])


def test_forest_that_does_not_fit_a_key_is_reported(native_lib):
    """More thresholds on a tabulated feature than a code byte holds: the family is reported as not memoisable (its
    requests are then always walked) instead of being keyed wrongly."""
    n = 300                                           # 300 distinct thresholds on seconds_remaining
    feat, thr, left, right, val, roots = [], [], [], [], [], []
    for t in range(n):
        roots.append(len(feat))
        feat += [6 + 5, -1, -1]; thr += [float(10 * t + 0.5), 0.0, 0.0]
        left += [len(left) + 1, -1, -1]; right += [len(right) + 2, -1, -1]; val += [0.0, 1.0, -1.0]
    f = art.Forest(
        name="wide", kind=art.KIND_XGB, link=art.LINK_SIGMOID, n_outputs=1, n_features=23, num_base=6, n_num=17,
        zero_is_missing=False, base_margin=np.zeros(1), scale=1.0, groups=[art.OneHotGroup("player", 0, list("abcdeU"))],
        feat=np.asarray(feat, np.int32), thr=np.asarray(thr, np.float32), left=np.asarray(left, np.int32),
        right=np.asarray(right, np.int32), default_left=np.zeros(len(feat), np.uint8), value=np.asarray(val, np.float64),
        tree_root=np.asarray(roots, np.int32), tree_out=np.zeros(n, np.int32))
    keys, meta = native.memo_keys_host(f, 0, np.array([[1, 10.0, 75.0, 0, 3600]], float), cols=(-1, -1), fold_values=np.zeros(17))
    assert keys is None and not meta["memoisable"] and "threshold" in meta["why"]
