"""Player mode of the CUDA engine (usage tables, dynamic one-hots, per-game player box) through the C-ABI:
against the fixtures produced by the unmodified reference (tests/golden/ref_players.npz) and against the
C oracle on injected and Philox draws."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN
from fast_monte_carlo_b200 import priors, usage
from fast_monte_carlo_b200.engine import Engine, MatchupSpec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def contexts(models_s2):
    focus = usage.build_focus_usage_tables(os.path.join(GOLDEN, "players_focus.csv"))
    sp = priors.load_sp_flex(priors.packaged_priors_path())
    out = {}
    for name in ("Kansas State", "Iowa State", "UTSA", "Ohio State", "Texas"):
        tc = priors.build_team_context_from_sp_flex(name, 2025, 1, sp, focus=focus, usage_dir=GOLDEN)
        out[name] = (tc, usage.resolve_team(tc, models_s2))
    return out


def _spec(contexts, a, b, n, g0=0, out_offset=0):
    (ta, ua), (tb, ub) = contexts[a], contexts[b]
    return MatchupSpec(a, b, ta.sp, tb.sp, n, g0, g0 + n, out_offset, usage=(ua, ub))


def _box_equal_philox(got, ref):
    """Philox runs: counts exact; yards within float tolerance (the state machine is bit-exact, but a normal
    draw in the AS241 tail goes through `log`, where the device and glibc results may differ in the last
    bit -- it vanishes in `ytg - yards` yet stays visible in a running sum of yards).  Injected-stream runs
    compare the float64 bits (test_players_injected_stream_vs_oracle)."""
    assert np.array_equal(got[..., 1:], ref[..., 1:])
    assert np.allclose(got[..., 0], ref[..., 0], rtol=1e-12, atol=1e-9)


def _frame(rows, cols):
    df = pd.DataFrame(rows, columns=cols)
    return df.sort_values(["sim", "team", "role", "player"], kind="stable").reset_index(drop=True)


def test_reference_golden_players(engine, contexts):
    """The reference's OWN trajectories and players table under a synthetic focus sheet."""
    t = np.load(os.path.join(GOLDEN, "ref_players.npz"))
    meta = json.loads(str(t["meta"]))
    cols = json.loads(str(t["player_cols"]))
    from oracle import c_oracle as co
    stream = co.make_stream(len(meta), int(t["stream_seed"]))
    frames = []
    for g, m in enumerate(meta):
        spec = _spec(contexts, m["team_a"], m["team_b"], 1, g0=g)
        engine.set_matchups([spec])
        r = engine.simulate_host(0, stream=stream[g:g + 1], want_trace=True, want_iters=True, want_players=True)
        k = int(t["iters"][g])
        assert r["iters"][0] == k, g
        assert np.array_equal(r["trace"][0, :k], t["traces"][g, :k]), g
        f = g & 1
        assert (r["scores"][0, f], r["scores"][0, f ^ 1]) == tuple(t["scores"][g])
        frames.append(usage.player_rows(r["players"], g, (m["team_a"], m["team_b"]), spec.usage))
    got = _frame(pd.concat(frames, ignore_index=True).values.tolist(), cols)
    want = _frame(json.loads(str(t["player_rows"])), cols)
    assert len(got) == len(want) > 50
    for c in cols:
        assert got[c].tolist() == want[c].tolist(), c


@pytest.mark.parametrize("a,b", [("Kansas State", "Iowa State"), ("UTSA", "Iowa State"), ("Ohio State", "Kansas State"),
                                 ("Ohio State", "UTSA")])
def test_players_injected_stream_vs_oracle(engine, oracle, models_s2, contexts, a, b):
    """Ohio State: usage from the fallback files (21 rushers, 12 targets, most of them unknown to every model)."""
    n = 4096
    stream = oracle.make_stream(n, 31)
    spec = _spec(contexts, a, b, n)
    engine.set_matchups([spec])
    got = engine.simulate_host(0, stream=stream, want_trace=True, want_iters=True, want_players=True)
    cfg = oracle.make_config(models_s2, spec.sp_a, spec.sp_b)
    ref = oracle.simulate(cfg, n, stream=stream, trace=True, usage=oracle.make_usage(spec.usage), n_slots=engine.n_slots)
    assert np.array_equal(got["scores"], ref["scores"])
    assert np.array_equal(got["iters"], ref["iters"])
    t0, t1 = got["trace"], ref["trace"]
    assert bool(((t0 == t1) | (np.isnan(t0) & np.isnan(t1))).all())
    assert np.array_equal(got["players"].dense(), ref["players"])          # float64 yards bit for bit
    if engine.n_slots:
        assert got["players"].dense()[..., 1].sum() > n                      # something was tracked
    else:                                                                    # usage files only: names matter, no box
        assert got["players"].dense().shape == (n, 2, 0, 6) and engine.ctx.has_usage
    assert len(contexts["Ohio State"][1].role["rush"].names) == 21 and not contexts["Ohio State"][1].slots


def test_players_philox_vs_oracle_booster(oracle, models_s2, contexts):
    """Philox on both sides, stage-2 booster on (its passer / target one-hots become dynamic rows too)."""
    n = 30000
    e = Engine(models_s2, device=0, stage2="booster")
    try:
        spec = _spec(contexts, "Kansas State", "Iowa State", n)
        e.set_matchups([spec])
        got = e.simulate_host(99, want_iters=True, want_players=True)
        cfg = oracle.make_config(models_s2, spec.sp_a, spec.sp_b, stage2="booster")
        ref = oracle.simulate(cfg, n, seed=99, usage=oracle.make_usage(spec.usage), n_slots=e.n_slots)
        assert np.array_equal(got["scores"], ref["scores"])
        assert np.array_equal(got["iters"], ref["iters"])
        _box_equal_philox(got["players"].dense(), ref["players"])
        for k in ("plays", "pass", "comp", "inc", "int", "sack", "run", "td"):
            assert got["counters"][k] == ref["counters"][k], k
    finally:
        e.close()


def test_trivial_usage_is_the_shipped_configuration(engine, contexts):
    n = 8192
    (ta, _), (tb, _) = contexts["UTSA"], contexts["Texas"]
    engine.set_matchups([MatchupSpec("UTSA", "Texas", ta.sp, tb.sp, n, 0, n, 0)])
    plain = engine.simulate_host(5)
    engine.set_matchups([_spec(contexts, "UTSA", "Texas", n)])
    assert not engine.ctx.has_usage           # both teams trivial: the production kernel runs
    again = engine.simulate_host(5)
    assert np.array_equal(plain["scores"], again["scores"])


def test_players_slate_two_matchups(engine, oracle, models_s2, contexts):
    """Two matchups in one launch, different usage tables, game-id slices as a rank would get them."""
    specs = [_spec(contexts, "Kansas State", "Iowa State", 3000, g0=1000, out_offset=0),
             _spec(contexts, "Iowa State", "UTSA", 2000, g0=0, out_offset=3000)]
    engine.set_matchups(specs)
    got = engine.simulate_host(7, want_players=True)
    for mi, s in enumerate(specs):
        cfg = oracle.make_config(models_s2, s.sp_a, s.sp_b)
        ref = oracle.simulate(cfg, s.games, game0=s.game_begin, matchup=mi, seed=7,
                              usage=oracle.make_usage(s.usage), n_slots=engine.n_slots)
        sl = slice(s.out_offset, s.out_offset + s.games)
        assert np.array_equal(got["scores"][sl], ref["scores"])
        _box_equal_philox(got["players"].dense()[sl], ref["players"])


def test_players_full_size_properties(engine, contexts):
    """4 M games in player mode (device buffers): size-independent invariants of the per-game box.
    Iowa State's sheet has one passer with share 1 and four receivers that sum to 1 (all tracked, no `__Other__`),
    so in every game its passer's completions / touchdowns / yards equal the sums over its receivers, and every
    pass call (attempt or sack) is exactly one target.  A launch over [0, N) equals two launches over the halves."""
    import torch
    n = 4_000_000
    spec = _spec(contexts, "Kansas State", "Iowa State", n)
    engine.set_matchups([spec])
    S = engine.n_slots
    dev = torch.device("cuda", 0)
    box = torch.zeros((n, 2, S, 2), dtype=torch.int64, device=dev)
    cnt = torch.zeros(32, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()
    engine.ctx.simulate_device(seed=77, counters=cnt.data_ptr(), players=box.data_ptr(), cuda_stream=st.cuda_stream)
    torch.cuda.synchronize()
    assert int(cnt[0]) == n
    yds = box[..., 0].view(torch.float64)
    c = box[..., 1]
    f = lambda k: (c >> (10 * k)) & 0x3FF                     # att|tgt, comp|rec, td, INT, sacks
    isu = contexts["Iowa State"][1]
    qb = [s for s, (r, _) in enumerate(isu.slots) if r == "pass"]
    wr = [s for s, (r, _) in enumerate(isu.slots) if r == "rec"]
    assert len(qb) == 1 and len(wr) == 4
    q = qb[0]
    assert torch.equal(f(1)[:, 1, q], f(1)[:, 1, wr].sum(dim=1))            # completions == receptions
    assert torch.equal(f(2)[:, 1, q], f(2)[:, 1, wr].sum(dim=1))            # passing TDs == receiving TDs
    assert torch.equal(f(0)[:, 1, q] + f(4)[:, 1, q], f(0)[:, 1, wr].sum(dim=1))   # attempts + sacks == targets
    assert torch.allclose(yds[:, 1, q], yds[:, 1, wr].sum(dim=1), rtol=0, atol=1e-9)
    assert bool((f(1) <= f(0)).all())                                       # completions <= attempts, receptions <= targets
    assert bool((f(3) <= f(0)).all())                                       # interceptions are attempts
    for t, tu in enumerate(spec.usage):
        for s_, (r, _) in enumerate(tu.slots):
            td_cap = f(0)[:, t, s_] if r == "rush" else f(1)[:, t, s_]      # a TD needs a carry / a completion
            assert bool((f(2)[:, t, s_] <= td_cap).all()), (t, s_)
    names = [isu.slots[s][1] for s in wr]
    assert int(f(0)[:, 1, wr[names.index("Avery Morrow")]].sum()) == 0      # share 0 (NaN usage): tracked, never sampled
    want = np.asarray(isu.role["rec"].share)[[isu.role["rec"].names.index(nm) for nm in names]]
    got = (f(0)[:, 1, wr].sum(dim=0).double() / f(0)[:, 1, wr].sum()).cpu().numpy()
    assert np.allclose(got, want / want.sum(), atol=2e-3)                   # targets follow the usage shares
    # sharding invariance (what a rank gets): two launches over the halves
    half = n // 2
    parts = []
    for g0 in (0, half):
        engine.set_matchups([_spec(contexts, "Kansas State", "Iowa State", half, g0=g0)])
        b = torch.zeros((half, 2, S, 2), dtype=torch.int64, device=dev)
        engine.ctx.simulate_device(seed=77, players=b.data_ptr(), cuda_stream=st.cuda_stream)
        torch.cuda.synchronize()
        parts.append(b)
    assert torch.equal(torch.cat(parts, dim=0), box)


def test_player_histograms_equal_box(engine, contexts):
    """The device's per-player histograms (atomics at game end, yards rounded to tenths with the fma midpoint test)
    equal the histograms the host builds from the per-game box with Python's round(x, 1)."""
    n = 200_000
    spec = _spec(contexts, "Kansas State", "Iowa State", n)
    engine.set_matchups([spec])
    r = engine.simulate_host(123, want_scores=False, want_hist=False, want_players=True, want_player_hist=True)
    assert r["counters"]["ph_overflow"] == 0
    want = usage.player_hist_from_box(r["players"], spec.usage)
    got = r["player_hist"][0]
    assert got.shape == want.shape == (2, engine.n_slots, usage.PH_BINS)
    assert np.array_equal(got, want)
    assert int(got[:, :, :usage.PH_YDS_BINS].sum()) > 10 * n
    # prop odds from the histograms == from the rows
    names = ("Kansas State", "Iowa State")
    a = usage.player_prop_odds_from_box(r["players"], names, spec.usage, "Kansas State", "Blake Watson", "rush_yards", 45.5)
    b = usage.player_prop_odds_from_hist(got, names, spec.usage, "Kansas State", "Blake Watson", "rush_yards", 45.5)
    for k, v in a.items():
        assert (abs(b[k] - v) < 1e-9) if isinstance(v, float) else (b[k] == v), k


def test_simulate_slate_with_players(engine, contexts):
    """api.simulate_slate in player mode: per-matchup player histograms (one launch for the slate) equal the
    single-matchup runs; the prop sheet is priced from them."""
    from fast_monte_carlo_b200 import api
    sheet = os.path.join(GOLDEN, "players_focus.csv")
    pairs = [("Kansas State", "Iowa State"), ("Ohio State", "Kansas State"), ("UTSA", "Texas")]
    res = api.simulate_slate(pairs, n=1500, seed=17, engine=engine, focus_csv=sheet, usage_dir=GOLDEN)
    assert res["_counters"]["games"] == 3 * 3000 and res["_counters"]["ph_overflow"] == 0
    S = res[pairs[0]]["player_hist"].shape[1]
    for m, (a, b) in enumerate(pairs):
        e = res[(a, b)]
        assert e["player_hist"].shape == (2, S, usage.PH_BINS) and e["games"] == 3000
        spec = _spec(contexts, a, b, 3000)
        if spec.usage[0].trivial and spec.usage[1].trivial:
            assert int(e["player_hist"].sum()) == 0
            continue
        # same matchup index as in the slate (it is part of the Philox counter): m empty matchups in front
        engine.set_matchups([MatchupSpec("pad", "pad", spec.sp_a, spec.sp_b, 0, 0, 0, 0, usage=spec.usage)] * m + [spec])
        single = engine.simulate_host(17, want_scores=False, want_hist=False, want_player_hist=True)["player_hist"][m]
        k = single.shape[1]
        assert np.array_equal(e["player_hist"][:, :k].astype(np.uint32), single)
    props = res[("Kansas State", "Iowa State")]["props"]
    assert len(props) >= 10 and {"Taylen Green", "Carson Hansen"} <= set(props["player"])
    assert len(res[("UTSA", "Texas")]["props"]) == 0                    # nobody tracked there


def test_set_usage_error_paths(engine, contexts, models_s2):
    """The C ABI's own checks (fmc_set_usage / fmc_simulate): capacity, order of calls, outputs without usage."""
    import copy
    from fast_monte_carlo_b200 import native
    spec = _spec(contexts, "Kansas State", "Iowa State", 100)
    engine.set_matchups([spec])
    bad = copy.deepcopy(spec.usage[0])
    g = models_s2["pass_stage1"].group("passer_name")
    ru = bad.role["pass"]
    ru.names = [g.categories[i] for i in range(5)]           # five passers the models know: one name row too many
    ru.share = np.full(5, 0.2)
    ru.slot = [-1] * 5
    ru.col = {m: [models_s2[m].group("passer_name").column_of(nm) if m != "run_yards" else -1 for nm in ru.names]
              for m in ru.col}
    with pytest.raises(native.FmcError, match="names of one role"):
        engine.ctx.set_usage([(bad, spec.usage[1])], 8)
    with pytest.raises(native.FmcError, match="n_matchups differs"):
        engine.ctx.set_usage([spec.usage, spec.usage], 8)
    (ta, _), (tb, _) = contexts["UTSA"], contexts["Texas"]
    engine.set_matchups([MatchupSpec("UTSA", "Texas", ta.sp, tb.sp, 10, 0, 10, 0)])
    with pytest.raises(native.FmcError, match="set_usage"):
        engine.simulate_host(1, want_players=True)


def test_reference_golden_players_adversarial(engine, contexts):
    """The reference's OWN player-mode games under adversarial draws (first / last usage entry, skipped zero-share
    entries, +-4 sigma yardage, 112-112 shoot-outs)."""
    from streams import adversarial_stream
    t = np.load(os.path.join(GOLDEN, "ref_players.npz"))
    meta = json.loads(str(t["adv_meta"]))
    cols = json.loads(str(t["player_cols"]))
    stream = adversarial_stream(len(meta), int(t["adv_stream_seed"]))
    frames = []
    for g, m in enumerate(meta):
        spec = _spec(contexts, m["team_a"], m["team_b"], 1, g0=g)
        engine.set_matchups([spec])
        r = engine.simulate_host(0, stream=stream[g:g + 1], want_trace=True, want_iters=True, want_players=True)
        k = int(t["adv_iters"][g])
        assert r["iters"][0] == k, (g, m["pattern"])
        assert np.array_equal(r["trace"][0, :k], t["adv_traces"][g, :k]), (g, m["pattern"])
        f = g & 1
        assert (r["scores"][0, f], r["scores"][0, f ^ 1]) == tuple(t["adv_scores"][g])
        frames.append(usage.player_rows(r["players"], g, (m["team_a"], m["team_b"]), spec.usage))
    got = _frame(pd.concat(frames, ignore_index=True).values.tolist(), cols)
    want = _frame(json.loads(str(t["adv_player_rows"])), cols)
    assert len(got) == len(want) > 10
    for c in cols:
        assert got[c].tolist() == want[c].tolist(), c
