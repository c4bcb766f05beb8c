"""The exact rank-keyed memo (csrc/fmc_memo.hpp, fmc_set_memo) must never change a result: every run here is compared
bit for bit with the unmemoised kernel and / or the CPU oracle."""
import numpy as np
import pytest

from conftest import ISU, KSU
from fast_monte_carlo_b200.engine import Engine, MatchupSpec

pytestmark = pytest.mark.gpu

KEYS = ("games", "plays", "iters", "pass", "comp", "inc", "int", "sack", "run", "td", "fga", "fg", "punt", "go")


def _run(models_s2, n, seed, spec=None, **kw):
    kw.setdefault("stage2", "booster")
    e = Engine(models_s2, device=0, **kw)
    try:
        e.set_matchups(spec or [MatchupSpec("A", "B", KSU, ISU, n, 0, n, 0)])
        return e.simulate_host(seed, want_iters=True)
    finally:
        e.close()


def test_memo_modes_bit_identical(models_s2, oracle):
    n = 300_000
    off = _run(models_s2, n, 20251018, memo="off")
    on = _run(models_s2, n, 20251018, memo="on")
    assert np.array_equal(on["scores"], off["scores"]) and np.array_equal(on["iters"], off["iters"])
    assert np.array_equal(on["hist"], off["hist"])
    for k in KEYS:
        assert on["counters"][k] == off["counters"][k], k
    c = on["counters"]
    assert off["counters"]["memo_probes"] == 0 and off["counters"]["memo_hits"] == 0
    assert c["memo_hits"] > 0.5 * c["memo_probes"] > 0
    # every request is either answered by the memo or walked
    total_requests = c["pass"] + (c["pass"] - c["comp"]) + c["comp"] + c["run"] + c["sack"]
    assert c["memo_probes"] == total_requests
    assert c["memo_hits"] + c["requests"] == total_requests
    assert c["memo_hits"] == sum(c[k] for k in ("memo_hits_s1", "memo_hits_s2", "memo_hits_pq", "memo_hits_rq", "memo_hits_sq", "memo_hits_pm"))
    assert c["memo_hits_pq"] > 0.9 * c["comp"] and c["memo_hits_pm"] == 0
    assert off["counters"]["requests"] == total_requests
    ref = oracle.simulate(oracle.make_config(models_s2, KSU, ISU, stage2="booster"), 20_000, seed=20251018)
    assert np.array_equal(on["scores"][:20_000], ref["scores"])


def test_memo_under_eviction_pressure(models_s2):
    """A table far too small for the run (direct-mapped, overwritten on conflict): still exact, just fewer hits."""
    n = 100_000
    off = _run(models_s2, n, 7, memo="off")
    tiny = _run(models_s2, n, 7, memo="on", memo_bytes=1 << 20)
    roomy = _run(models_s2, n, 7, memo="on")
    assert np.array_equal(tiny["scores"], off["scores"]) and np.array_equal(tiny["iters"], off["iters"])
    assert np.array_equal(roomy["scores"], off["scores"])
    assert 0 < tiny["counters"]["memo_hits"] < roomy["counters"]["memo_hits"]


def test_persistent_memo_and_table_change(models_s2, oracle):
    """mode 2 keeps the table between launches on the same tables (second launch: nearly every request hits) and drops
    it when the tables change."""
    n = 60_000
    e = Engine(models_s2, device=0, stage2="booster", memo="persistent")
    try:
        e.set_matchups([MatchupSpec("A", "B", KSU, ISU, n, 0, n, 0)])
        a = e.simulate_host(11, want_iters=True)
        b = e.simulate_host(11, want_iters=True)
        assert np.array_equal(a["scores"], b["scores"]) and np.array_equal(a["iters"], b["iters"])
        assert b["counters"]["memo_hits"] > a["counters"]["memo_hits"]
        assert b["counters"]["memo_hits"] > 0.85 * b["counters"]["memo_probes"]      # direct mapped: later inserts overwrite a fifth of the entries
        c = e.simulate_host(12)                      # another seed on the warm table
        ref = oracle.simulate(oracle.make_config(models_s2, KSU, ISU, stage2="booster"), 4000, seed=12)
        assert np.array_equal(c["scores"][:4000], ref["scores"])
        # the same ranges again through set_matchups: tables (and the memo) are kept
        e.set_matchups([MatchupSpec("A", "B", KSU, ISU, n, 0, n, 0)])
        d = e.simulate_host(11)
        assert np.array_equal(d["scores"], a["scores"]) and d["counters"]["memo_hits"] > 0.85 * d["counters"]["memo_probes"]
        # another pair: new tables, the old entries must not be served
        other = ((0.0, 28.0, 27.5), (31.7, 41.9, 10.1))
        e.set_matchups([MatchupSpec("C", "D", other[0], other[1], n, 0, n, 0)])
        g = e.simulate_host(11)
        ref = oracle.simulate(oracle.make_config(models_s2, other[0], other[1], stage2="booster"), 4000, seed=11)
        assert np.array_equal(g["scores"][:4000], ref["scores"])
        assert g["counters"]["memo_hits"] < 0.8 * g["counters"]["memo_probes"]
    finally:
        e.close()


def test_memo_slate_keys_carry_the_matchup(models_s2, oracle):
    """Several matchups share one table: the key carries the matchup, so entries never cross."""
    pairs = [(KSU, ISU), (ISU, KSU), ((0.0, 28.0, 27.5), (31.7, 41.9, 10.1)), ((-19.3, 17.3, 36.6), (27.9, 40.4, 12.6))]
    games = 6000
    spec, off_ = [], 0
    for a, b in pairs:
        spec.append(MatchupSpec("a", "b", a, b, games, 0, games, off_))
        off_ += games
    on = _run(models_s2, 0, 5, spec=spec, memo="on")
    assert on["counters"]["memo_hits"] > 0
    for m, (a, b) in enumerate(pairs):
        ref = oracle.simulate(oracle.make_config(models_s2, a, b, stage2="booster"), games, matchup=m, seed=5)
        assert np.array_equal(on["scores"][m * games:(m + 1) * games], ref["scores"]), m


@pytest.mark.parametrize("variant", [
    dict(policy="play_model", stage2="booster", sampler="quantile_interp", play_temp=1.3, qy_noise=0.7),
    dict(stage2="standin"),
])
def test_memo_with_every_model_family(models_s2, variant):
    """play_model.xgb policy (its own memo region), quantile-interpolation sampler, stage-2 stand-in."""
    n = 40_000
    off = _run(models_s2, n, 3, memo="off", **variant)
    on = _run(models_s2, n, 3, memo="on", **variant)
    assert np.array_equal(on["scores"], off["scores"]) and np.array_equal(on["iters"], off["iters"])
    assert on["counters"]["memo_hits"] > 0
