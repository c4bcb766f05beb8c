"""GPU tree-ensemble evaluator vs the oracle, through the C-ABI (fmc_tree_predict_host).
Tolerance: the north star asks for 1e-5 relative on raw margins; the kernels accumulate in the same
precision and order as the oracle, so the tests demand bit equality."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from fast_monte_carlo_b200 import artifacts as art
from oracle import tree_oracle as to
from test_pack import _rows

pytestmark = pytest.mark.gpu

ALL = ["pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards", "run_fumble", "play_model"]


def _oracle_margins(oracle, f, name, rows, cols, tb=0, te=-1):
    x = to.play_model_features(f, rows[:, :12]) if f.scaler_cols is not None else rows
    return oracle.predict(name, x, np.tile(np.array(cols, dtype=np.int32), (rows.shape[0], 1)), f.n_outputs, tb, te)


@pytest.mark.parametrize("name", ALL)
def test_margins_bit_exact(engine, oracle, models_s2, name):
    f = models_s2[name]
    rows = _rows(20000, 11)          # includes exact-zero distance / ytg / score_diff / SP rating and down >= 5
    cols = [g.column_of("Unknown") for g in f.groups if g.name != "coach"] + [-1, -1]
    ref = _oracle_margins(oracle, f, name, rows, cols[:2])
    got = engine.predict(name, rows[:, :f.n_num])
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12)
    assert rel.max() <= 1e-5


def test_sklearn_golden_through_gpu(engine, models_s2):
    """The live sklearn pipelines' own predictions (all-"Unknown" rows of the golden file)."""
    g = np.load(os.path.join(GOLDEN, "sklearn_quantiles.npz"))
    for fam in ("pass_yards", "run_yards", "sack_yards"):
        f = models_s2[fam]
        cols = [gr.column_of("Unknown") for gr in f.groups] + [-1]
        act = g[f"{fam}/active"]
        sel = np.all(act == np.array(cols[:2]), axis=1)
        assert sel.sum() > 100
        got = engine.predict(fam, g["num"][sel])
        assert np.array_equal(got, g[f"{fam}/pred"][sel])


def test_per_row_names_all_golden_rows(engine, oracle, models_s2):
    """Every row of the golden file with ITS OWN passer / target / rusher columns (FMC:744, 756, 784-809 predict rows
    with arbitrary names): fmc_tree_predict_cols_host against the live sklearn pipelines' predictions, bit for bit,
    and against the oracle for the boosters."""
    g = np.load(os.path.join(GOLDEN, "sklearn_quantiles.npz"))
    for fam in ("pass_yards", "run_yards", "sack_yards"):
        act = g[f"{fam}/active"].astype(np.int32)
        assert len(np.unique(act, axis=0)) > 5                     # many different name pairs in one call
        got = engine.predict(fam, g["num"], hot_cols=act)
        assert np.array_equal(got, g[f"{fam}/pred"]), fam
    rng = np.random.default_rng(8)
    for name in ("pass_stage1", "pass_stage2", "run_fumble"):
        f = models_s2[name]
        rows = _rows(3000, 17)
        cols = np.full((rows.shape[0], 2), -1, dtype=np.int32)
        for gi, grp in enumerate(f.groups[:2]):
            pick = rng.integers(-1, len(grp.categories), rows.shape[0])
            cols[:, gi] = np.where(pick < 0, -1, grp.base + pick)
        ref = oracle.predict(name, rows, cols, f.n_outputs)
        got = engine.predict(name, rows[:, :f.n_num], hot_cols=cols)
        assert np.array_equal(got, ref), name
    # by name, through the category lists
    f = models_s2["pass_yards"]                                    # passer_name + target_name
    rows = _rows(64, 3)
    names = [("Adrian Martinez", "Aaron Anderson") if i % 2 else ("nobody at all", None) for i in range(64)]
    cols = np.array([[f.groups[0].column_of(a), f.groups[1].column_of(b)] for a, b in names], dtype=np.int32)
    assert cols.max() > 0 and cols.min() == -1
    assert np.array_equal(engine.predict("pass_yards", rows, names=names), oracle.predict("pass_yards", rows, cols, 3))


def test_nan_is_missing_for_xgboost(engine, oracle, models_s2):
    """Booster.predict on a DataFrame row treats NaN as missing (the node's default branch): honoured for every numeric
    that is not a 0/1 flag; scikit-learn pipelines raise on NaN."""
    from fast_monte_carlo_b200.native import FmcError
    rng = np.random.default_rng(4)
    for name in ("pass_stage1", "pass_stage2", "run_fumble", "play_model"):
        f = models_s2[name]
        rows = _rows(4000, 31)
        nonflag = [k for k in range(f.n_num) if k not in (3, 12, 13, 14, 16)]
        mask = rng.random((rows.shape[0], len(nonflag))) < 0.15
        sub = rows[:, nonflag]
        sub[mask] = np.nan
        rows[:, nonflag] = sub
        cols = ([g.column_of("Unknown") for g in f.groups if g.name != "coach"] + [-1, -1])[:2]
        ref = _oracle_margins(oracle, f, name, rows, cols)
        got = engine.predict(name, rows[:, :f.n_num])
        assert np.array_equal(got, ref), name
        clean = _oracle_margins(oracle, f, name, np.nan_to_num(rows, nan=1.0), cols)
        assert not np.array_equal(ref, clean)                  # the NaNs did take other branches
    bad = _rows(8, 1)
    bad[3, 2] = np.nan
    with pytest.raises(FmcError, match="NaN"):
        engine.predict("pass_yards", bad)
    bad = _rows(8, 1)
    bad[2, 3] = np.nan                                         # is_red_zone of the dense-fed play model
    with pytest.raises(FmcError, match="flag"):
        engine.predict("play_model", bad[:, :12])


def test_named_players_and_tree_ranges(engine, oracle, models_s2):
    """Real one-hot columns (fmc_set_active_columns) and iteration_range (sim_helpers.py:22-23)."""
    p = json.load(open(os.path.join(GOLDEN, "xgb_provisional.json")))
    f = models_s2["pass_stage1"]
    mid = art.MODEL_IDS["pass_stage1"]
    col = f.groups[0].column_of("Caleb Williams")
    engine.ctx.set_active_columns(mid, col, -1)
    try:
        r0 = np.array([p["r0"]], dtype=float)
        for v in p["stage1"][:2]:
            got = engine.ctx.tree_predict_host(mid, r0, 1, 0, v["trees"])
            assert abs(got[0, 0] - v["margin"]) < 2e-7
        rows = _rows(3000, 12)
        got = engine.ctx.tree_predict_host(mid, rows, 1, 0, 68)
        assert np.array_equal(got, _oracle_margins(oracle, f, "pass_stage1", rows, (col, -1), 0, 68))
    finally:
        engine.ctx.set_active_columns(mid, -1, -1)
    g = np.load(os.path.join(GOLDEN, "sklearn_quantiles.npz"))
    fq = models_s2["pass_yards"]
    act = g["pass_yards/active"]
    combos, counts = np.unique(act, axis=0, return_counts=True)
    pick = combos[np.argsort(-counts)[1]]          # most common non-"Unknown,Unknown" combination
    sel = np.all(act == pick, axis=1)
    engine.ctx.set_active_columns(art.MODEL_IDS["pass_yards"], int(pick[0]), int(pick[1]))
    try:
        got = engine.predict("pass_yards", g["num"][sel])
        assert np.array_equal(got, g["pass_yards/pred"][sel])
    finally:
        engine.ctx.set_active_columns(art.MODEL_IDS["pass_yards"], fq.groups[0].column_of("Unknown"),
                                      fq.groups[1].column_of("Unknown"))


def test_play_model_coach_column(engine, oracle, models_s2):
    f = models_s2["play_model"]
    col = f.group("coach").column_of("Chris Klieman")
    assert col >= 12
    rows = _rows(2000, 13)
    ref = _oracle_margins(oracle, f, "play_model", rows, (col, -1))
    got = engine.predict("play_model", rows[:, :12], coach="Chris Klieman")
    assert np.array_equal(got, ref)
    assert not np.array_equal(got, engine.predict("play_model", rows[:, :12]))


def test_empty_and_ragged_sizes(engine, oracle, models_s2):
    f = models_s2["run_yards"]
    cols = [f.groups[0].column_of("Unknown"), -1]
    assert engine.predict("run_yards", np.zeros((0, 17))).shape == (0, 3)
    for n in (1, 31, 33, 255, 257, 1000):
        rows = _rows(n, n)
        assert np.array_equal(engine.predict("run_yards", rows), _oracle_margins(oracle, f, "run_yards", rows, cols))


def test_forest_larger_than_one_window(models_s2, native_lib):
    """A node table above 1 MiB spans several windows (per-group window index in the root stream)."""
    from fast_monte_carlo_b200 import synth
    from fast_monte_carlo_b200.engine import Engine
    big = synth.synthetic_stage2(models_s2, seed=5, rounds=1500)
    forests = dict(models_s2.forests)
    forests["pass_stage2"] = big
    e = Engine(art.ModelSet(forests, source="big-stage2"), device=0, stage2="standin")
    try:
        rows = _rows(4000, 21)
        cols = [g.column_of("Unknown") for g in big.groups]
        got = e.predict("pass_stage2", rows)
        ref = to.raw_margin(big, rows, np.tile(np.array(cols[:2]), (rows.shape[0], 1)))
        assert np.array_equal(got, ref)
    finally:
        e.close()
