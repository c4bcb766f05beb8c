"""One-hot / numerics layout and the scaler against the live sklearn transformers (golden)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from fast_monte_carlo_b200 import artifacts as art
from oracle import tree_oracle as to


@pytest.mark.parametrize("name", ["pass_stage1"])
def test_column_transformer_layout(models, name):
    g = np.load(os.path.join(GOLDEN, "transformers.npz"))
    f = models[name]
    num = g["num"]
    indptr, indices, data = g[f"{name}/indptr"], g[f"{name}/indices"], g[f"{name}/data"]
    assert tuple(g[f"{name}/shape"]) == (num.shape[0], f.n_features)
    for r in range(num.shape[0]):
        want = dict(zip(indices[indptr[r]:indptr[r + 1]].tolist(), data[indptr[r]:indptr[r + 1]].tolist()))
        mine = {}
        for grp in f.groups:
            col = grp.column_of(str(g[f"{name}/name/{grp.name}"][r]))
            if col >= 0:
                mine[col] = 1.0
        for k in range(art.N_NUM):
            if num[r, k] != 0.0:          # the CSR hstack drops exact zeros
                mine[f.num_base + k] = float(num[r, k])
        assert mine == want


def test_stage2_preprocessor_layout():
    """The stage-2 booster is missing but its preprocessor ships: 486 passers ("Unknown" at 467),
    one target category, numerics at 487 -- the layout synth.synthetic_stage2 assumes."""
    g = np.load(os.path.join(GOLDEN, "transformers.npz"))
    assert tuple(g["pass_stage2/groupbase/passer_name"]) == (0, 486)
    assert tuple(g["pass_stage2/groupbase/target_name"]) == (486, 1)
    assert int(g["pass_stage2/shape"][1]) == 504
    names = g["pass_stage2/name/passer_name"]
    indptr, indices = g["pass_stage2/indptr"], g["pass_stage2/indices"]
    r = int(np.flatnonzero(names == "Unknown")[0])
    assert 467 in indices[indptr[r]:indptr[r + 1]]


def test_scaler_matches_live_standard_scaler(models):
    g = np.load(os.path.join(GOLDEN, "transformers.npz"))
    f = models["play_model"]
    raw12 = np.zeros((g["scaler/in"].shape[0], 12))
    raw12[:, f.scaler_cols] = g["scaler/in"]
    out = to.play_model_features(f, raw12)
    assert np.array_equal(out[:, f.scaler_cols], g["scaler/out"])


def test_forest_structure(models):
    """Gate G1: parsed-booster structural self-checks."""
    for name, f in models.forests.items():
        art.check_forest(f)
    s1 = models["pass_stage1"]
    assert (s1.n_trees, s1.n_nodes, s1.n_features, s1.num_base) == (188, 20348, 580, 563)
    assert s1.best_iteration == 67
    onehot = (s1.feat >= 0) & (s1.feat < s1.num_base)
    assert np.all(s1.thr[onehot] == np.float32(2.00001)) and not s1.default_left[onehot].any()
    pm = models["play_model"]
    assert (pm.n_trees, pm.n_nodes, pm.n_outputs, pm.n_features) == (1000, 50030, 5, 180)
    assert [models[k].n_trees for k in ("pass_yards", "run_yards", "sack_yards")] == [1200] * 3
    assert list(models["pass_yards"].base_margin) == [2.0, 9.0, 26.0]
    assert models["pass_yards"].group("passer_name").column_of("Unknown") == 491
    assert models["pass_yards"].group("target_name").column_of("Unknown") == 2877
    assert models["run_yards"].group("rusher_name").column_of("Unknown") == 1780
    assert models["sack_yards"].group("passer_name").column_of("Unknown") == 347
    assert s1.group("passer_name").column_of("Unknown") == -1


def test_threshold_floor_preserves_predicate():
    rng = np.random.default_rng(0)
    t64 = rng.normal(0, 50, 20000)
    t32 = art.f64_to_f32_floor(t64)
    x = np.concatenate([t32, np.nextafter(t32, np.float32(np.inf)), np.nextafter(t32, np.float32(-np.inf))])
    tt = np.concatenate([t64] * 3)
    t3 = np.concatenate([t32] * 3)
    assert np.array_equal(x.astype(np.float64) <= tt, x <= t3)
