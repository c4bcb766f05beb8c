"""The C++ specialiser/packer (fmc_pack.hpp) on CPU: the packed tables, walked in NumPy by the test,
must reproduce the oracle's margins on the ORIGINAL trees bit for bit."""
import numpy as np
import pytest

import packed_walk as pw
from fast_monte_carlo_b200 import native
from oracle import tree_oracle as to

KSU, ISU, UTSA = (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), (0.0, 28.0, 27.5)


def _rows(n, seed):
    rng = np.random.default_rng(seed)
    down = rng.choice([1, 2, 3, 4, 5, 6], size=n)
    dist = np.round(rng.uniform(0.3, 25, n), 2)
    ytg = np.round(rng.uniform(0.5, 102, n), 2)
    sd = rng.integers(-28, 29, n)
    sd[::4] = 0
    sec = rng.integers(1, 3601, n)
    num = np.zeros((n, 17))
    num[:, 0] = down; num[:, 1] = dist; num[:, 2] = ytg; num[:, 3] = ytg <= 20; num[:, 4] = sd; num[:, 5] = sec
    num[:, 6] = rng.integers(0, 4, n); num[:, 7] = rng.integers(0, 4, n)
    num[:, 8:12] = np.round(rng.normal(5, 15, (n, 4)), 1)
    num[::7, 8] = 0.0
    num[:, 12] = dist >= ytg - 0.5; num[:, 13] = (down == 4) & (dist <= 2); num[:, 14] = ytg <= 33
    num[:, 15] = np.where(sec > 1800, 1, 2); num[:, 16] = (sec % 1800) <= 120
    num[:3, 1] = 0.0
    num[3:6, 2] = 0.0
    return num


def _setup(models_s2, name, player="Unknown"):
    f = models_s2[name]
    cols = [g.column_of(player) for g in f.groups if g.name != "coach"] + [-1, -1]
    skl = f.kind == 1
    zm = bool(f.zero_is_missing) and not skl
    scaler = (f.scaler_cols, f.scaler_mean, f.scaler_scale) if f.scaler_cols is not None else None
    return f, cols[:2], skl, zm, scaler


@pytest.mark.parametrize("name", ["pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards", "run_fumble", "play_model"])
def test_predict_preset(models_s2, native_lib, name):
    f, cols, skl, zm, scaler = _setup(models_s2, name)
    num = _rows(96, 1)[:, :f.n_num]
    slots, stream, consts, meta = native.pack_forest_host(f, mode=1, cols=cols)
    got = pw.walk(slots, stream, consts, meta, pw.predict_rows(num, zm, scaler, xgb=not skl), skl, f.base_margin)
    x = to.play_model_features(f, num) if scaler else num
    ref = to.raw_margin(f, x, np.tile(np.array(cols), (num.shape[0], 1)))
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("name", ["pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards", "play_model"])
@pytest.mark.parametrize("off,de", [(KSU, ISU), (UTSA, KSU), (ISU, UTSA)])
def test_sim_preset_folds_orientation_constants(models_s2, native_lib, name, off, de):
    f, cols, skl, zm, scaler = _setup(models_s2, name)
    num = _rows(64, 2)[:, :f.n_num]
    fv = np.zeros(17)
    fv[6] = fv[7] = 3.0
    fv[8], fv[9], fv[10], fv[11] = off[0], off[1], de[2], de[0]
    num[:, 6:12] = fv[6:12]
    slots, stream, consts, meta = native.pack_forest_host(f, mode=0, cols=cols, fold_values=fv)
    got = pw.walk(slots, stream, consts, meta, pw.sim_rows(num, zm, scaler), skl, f.base_margin)
    x = to.play_model_features(f, num) if scaler else num
    ref = to.raw_margin(f, x, np.tile(np.array(cols), (num.shape[0], 1)))
    assert np.array_equal(got, ref)
    # specialisation must shrink the table (pass-through chains included), never grow it
    assert len(slots) <= f.n_nodes + 16


def test_tree_range_and_named_player(models_s2, native_lib):
    """iteration_range (best_iteration + 1 rounds, pass_outcome_infer.py:57) and a real passer column."""
    f, _, skl, zm, scaler = _setup(models_s2, "pass_stage1")
    col = f.groups[0].column_of("Caleb Williams")
    assert col == 82
    num = _rows(40, 3)
    slots, stream, consts, meta = native.pack_forest_host(f, mode=1, cols=(col, -1), tree_begin=0, tree_end=68)
    assert meta["rounds"] == 68
    got = pw.walk(slots, stream, consts, meta, pw.predict_rows(num, zm, None, xgb=not skl), skl, f.base_margin)
    ref = to.raw_margin(f, num, np.tile(np.array([col, -1]), (40, 1)), 0, 68)
    assert np.array_equal(got, ref)


def test_constant_trees_and_group_padding(models_s2, native_lib):
    """Trees that fold to one leaf leave the walk (they become ordered constants); the last group of an
    output is completed with padding trees that end in the +0.0 leaf."""
    f = models_s2["sack_yards"]
    fv = np.zeros(17)
    fv[6] = fv[7] = 3.0
    fv[8], fv[9], fv[10], fv[11] = KSU[0], KSU[1], ISU[2], ISU[0]
    cols = [g.column_of("Unknown") for g in f.groups] + [-1, -1]
    slots, stream, consts, meta = native.pack_forest_host(f, mode=0, cols=cols[:2], fold_values=fv)
    assert meta["rounds"] == 400 and meta["ilp"] == 4
    walked = 4 * sum(meta["n_groups"][:3])
    assert meta["constants"] > 0 and meta["constants"] + walked >= 1200 and walked < 1200
    assert len(stream) == walked
    # a model with nothing to fold keeps every tree
    f1 = models_s2["pass_yards"]
    cols1 = [g.column_of("Unknown") for g in f1.groups] + [-1, -1]
    _, stream1, _, meta1 = native.pack_forest_host(f1, mode=1, cols=cols1[:2])
    assert meta1["constants"] + len(stream1) >= 1200


def test_table_larger_than_one_window(models_s2, native_lib):
    """A forest whose node table exceeds 1 MiB spans several windows: every group stays inside one and
    carries its window index (fmc_pack.hpp)."""
    from fast_monte_carlo_b200 import synth
    big = synth.synthetic_stage2(models_s2, seed=5, rounds=1500)
    cols = [g.column_of("Unknown") for g in big.groups] + [-1, -1]
    slots, stream, consts, meta = native.pack_forest_host(big, mode=1, cols=cols[:2])
    assert len(slots) * 8 > (1 << 20)
    hi = (stream >> np.uint64(32)).astype(np.uint32).reshape(-1, 4)
    windows = (hi[:, 2] & 7) | ((hi[:, 3] & 7) << 3)
    assert windows.max() >= 1 and np.all(np.diff(windows[: meta["n_groups"][0]].astype(int)) >= 0)
    num = _rows(8, 9)
    got = pw.walk(slots, stream, consts, meta, pw.predict_rows(num, True, None), False, big.base_margin)
    ref = to.raw_margin(big, num, np.tile(np.array(cols[:2]), (num.shape[0], 1)))
    assert np.array_equal(got, ref)
