"""Player mode of the specialiser/packer on CPU: one-hot columns of the usage entries become per-request
0/1 feature rows (fmc_pack.hpp `dyn_col`), every other name column folds to zero.  The packed tables, walked
in NumPy, must equal the oracle's margins on the ORIGINAL trees for every sampled (passer, target) / rusher."""
import os

import numpy as np
import pytest

import packed_walk as pw
from conftest import GOLDEN
from fast_monte_carlo_b200 import native, priors, usage
from oracle import tree_oracle as to
from test_pack import _rows

DYN_ROW0, MAX_PASSERS, MAX_USAGE = 15, 4, 8      # kDynRow0, FMC_MAX_PASSER_ROWS, FMC_MAX_NAME_ROWS


@pytest.fixture(scope="module")
def teams(models_s2):
    focus = usage.build_focus_usage_tables(os.path.join(GOLDEN, "players_focus.csv"))
    sp = priors.load_sp_flex(priors.packaged_priors_path())
    out = {}
    for name in ("Kansas State", "Iowa State", "UTSA", "Ohio State"):
        tc = priors.build_team_context_from_sp_flex(name, 2025, 1, sp, focus=focus, usage_dir=GOLDEN)
        out[name] = (tc, usage.resolve_team(tc, models_s2))
    return out


@pytest.mark.parametrize("name", ["pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards"])
@pytest.mark.parametrize("off,de", [("Kansas State", "Iowa State"), ("Iowa State", "UTSA"), ("UTSA", "Kansas State"),
                                    ("Ohio State", "Kansas State")])
def test_dynamic_one_hot_rows(models_s2, native_lib, teams, name, off, de):
    f = models_s2[name]
    skl = f.kind == 1
    zm = bool(f.zero_is_missing) and not skl
    tc_o, tu = teams[off]
    tc_d, _ = teams[de]
    n = 96
    num = _rows(n, 4)[:, :f.n_num]
    fv = np.zeros(17)
    fv[6] = fv[7] = 3.0
    fv[8], fv[9], fv[10], fv[11] = tc_o.sp_rating, tc_o.sp_offense, tc_d.sp_defense, tc_d.sp_rating
    num[:, 6:12] = fv[6:12]
    rng = np.random.default_rng(7)
    dyn = {}
    if name == "run_yards":
        ru = tu.role["rush"]
        e0 = rng.integers(0, len(ru.names), n)
        cols = np.stack([np.asarray(ru.col[name])[e0], np.full(n, -1)], axis=1)
        nr = np.asarray(usage.name_rows(ru))
        hot = [(nr[e0], DYN_ROW0)]
        dyn.update({c: DYN_ROW0 + nr[e] for e, c in enumerate(ru.col[name]) if c >= 0})
    else:
        qb, wr = tu.role["pass"], tu.role["rec"]
        e0 = rng.integers(0, len(qb.names), n)
        e1 = rng.integers(0, len(wr.names), n)
        cols = np.stack([np.asarray(qb.col[name])[e0], np.asarray(wr.col[name])[e1]], axis=1)
        nq, nw = np.asarray(usage.name_rows(qb)), np.asarray(usage.name_rows(wr))
        hot = [(nq[e0], DYN_ROW0), (nw[e1], DYN_ROW0 + MAX_PASSERS)]
        dyn.update({c: DYN_ROW0 + nq[e] for e, c in enumerate(qb.col[name]) if c >= 0})
        dyn.update({c: DYN_ROW0 + MAX_PASSERS + nw[e] for e, c in enumerate(wr.col[name]) if c >= 0})
    slots, stream, consts, meta = native.pack_forest_host(f, mode=0, cols=(-1, -1), fold_values=fv, dyn=dyn)
    rows = np.zeros((n, DYN_ROW0 + MAX_PASSERS + MAX_USAGE), dtype=np.float32)
    rows[:, :DYN_ROW0] = pw.sim_rows(num, zm, None)
    for r, r0 in hot:            # a name without a name row lights nothing
        k = np.nonzero(r >= 0)[0]
        rows[k, r0 + r[k]] = 1.0
    got = pw.walk(slots, stream, consts, meta, rows, skl, f.base_margin)
    ref = to.raw_margin(f, num, cols)
    assert np.array_equal(got, ref)
    if off != "UTSA" and name in ("pass_stage1", "pass_yards", "run_yards", "sack_yards"):
        # the fixture's names are split on by these models: the rows must matter
        plain = to.raw_margin(f, num, np.full((n, 2), -1))
        assert not np.array_equal(ref, plain)
