"""GPU simulation kernel vs the reference's golden trajectories and vs the oracle, through the
C-ABI (fmc_simulate_host).  Integer trajectories must be bit-exact under an injected draw stream;
float64 state (distance, yardsToGoal) is compared bit-exact as well (stronger than the float
tolerance the north star allows)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ISU, KSU
from fast_monte_carlo_b200 import outputs
from fast_monte_carlo_b200.engine import Engine, MatchupSpec

pytestmark = pytest.mark.gpu

UTSA, OSU = (0.0, 28.0, 27.5), None


def _trace_equal(a, b):
    return bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def test_reference_golden_trajectories(engine):
    """The reference's OWN simulate_game states (tests/golden/ref_trajectories.npz)."""
    t = np.load(os.path.join(GOLDEN, "ref_trajectories.npz"))
    meta = json.loads(str(t["meta"]))
    n = len(meta)
    from oracle import c_oracle as co
    stream = co.make_stream(n, int(t["stream_seed"]))
    # games alternate between two matchups (pairs of games); run each matchup with the game ids it had
    for mi in range(2):
        idx = [g for g in range(n) if (g // 2) % 2 == mi]
        first = meta[idx[0]]
        spA, spB = first["sp_first"], first["sp_second"]
        for g in idx:
            eng_stream = stream[g:g + 1]
            engine.set_matchups([MatchupSpec("A", "B", tuple(spA), tuple(spB), 1, g, g + 1, 0)])
            r = engine.simulate_host(0, stream=eng_stream, want_trace=True, want_iters=True)
            k = int(t["iters"][g])
            assert r["iters"][0] == k
            assert np.array_equal(r["trace"][0, :k], t["traces"][g, :k])
            f = g & 1
            assert (r["scores"][0, f], r["scores"][0, f ^ 1]) == tuple(t["scores"][g])
            assert r["counters"]["plays"] == int(t["plays"][g])


def test_reference_golden_adversarial_trajectories(engine):
    """The reference's OWN states under adversarial draws (tests/golden/ref_trajectories_adversarial.npz)."""
    from streams import adversarial_stream
    t = np.load(os.path.join(GOLDEN, "ref_trajectories_adversarial.npz"))
    meta = json.loads(str(t["meta"]))
    stream = adversarial_stream(len(meta), int(t["stream_seed"]))
    for g, m in enumerate(meta):
        spA, spB = (m["sp_second"], m["sp_first"]) if g & 1 else (m["sp_first"], m["sp_second"])
        engine.set_matchups([MatchupSpec("A", "B", tuple(spA), tuple(spB), 1, g, g + 1, 0)])
        r = engine.simulate_host(0, stream=stream[g:g + 1], want_trace=True, want_iters=True)
        k = int(t["iters"][g])
        assert r["iters"][0] == k, (g, m["pattern"])
        assert np.array_equal(r["trace"][0, :k], t["traces"][g, :k]), (g, m["pattern"])
        f = g & 1
        assert (r["scores"][0, f], r["scores"][0, f ^ 1]) == tuple(t["scores"][g])
        assert r["counters"]["plays"] == int(t["plays"][g])


@pytest.mark.parametrize("spa,spb", [(KSU, ISU), ((0.0, 28.0, 27.5), (31.7, 41.9, 10.1))])
def test_injected_stream_bit_exact_16384(engine, oracle, models_s2, spa, spb):
    """BASELINE config 2's check: 16,384 games, [games, 360, 16] float64 draws from default_rng(7)."""
    n = 16384
    stream = oracle.make_stream(n, 7)
    engine.set_matchups([MatchupSpec("A", "B", spa, spb, n, 0, n, 0)])
    got = engine.simulate_host(0, stream=stream, want_trace=True, want_iters=True)
    ref = oracle.simulate(oracle.make_config(models_s2, spa, spb), n, stream=stream, trace=True)
    assert np.array_equal(got["scores"], ref["scores"])
    assert np.array_equal(got["iters"], ref["iters"])
    assert _trace_equal(got["trace"], ref["trace"])
    for k in ("plays", "iters", "pass", "comp", "inc", "int", "sack", "run", "td", "fga", "fg", "punt", "go"):
        assert got["counters"][k] == ref["counters"][k], k
    assert got["counters"]["games"] == n
    assert np.array_equal(got["hist"][0], outputs.histogram_from_scores(got["scores"]))


def test_philox_scores_equal_oracle(engine, oracle, models_s2):
    """Same Philox key/counter scheme on both sides: identical games (gate G6 at 200 k games here; the recorded
    1 M-game run of scripts/g6_full.py is under profiles/)."""
    n = 200_000
    engine.set_matchups([MatchupSpec("A", "B", KSU, ISU, n, 0, n, 0)])
    got = engine.simulate_host(20251018, want_iters=True)
    ref = oracle.simulate(oracle.make_config(models_s2, KSU, ISU), n, seed=20251018)
    assert np.array_equal(got["scores"], ref["scores"])
    assert np.array_equal(got["iters"], ref["iters"])
    assert got["counters"]["plays"] == ref["counters"]["plays"]
    other = engine.simulate_host(20251019)
    assert not np.array_equal(other["scores"], got["scores"])


def _engine_variant(models_s2, **kw):
    kw.setdefault("stage2", "standin")      # Engine's "auto" would pick the synthetic booster of models_s2
    return Engine(models_s2, device=0, **kw)


@pytest.mark.parametrize("memo", ["on", "off"])
def test_g6_full_size_digest(models_s2, memo):
    """Gate G6 at the full size of BASELINE configs[1] (GPU vs host twin on ALL 10 M Philox games): the C
    oracle played the 10,000,000 games of the headline workload once on host cores (scripts/g6_digest.py, 2.6 hours on 8
    cores) and recorded the SHA-1 of the score table and its event counters (tests/golden/g6_digest.json); the GPU plays
    the same games here in one call -- with the exact memo as shipped (`sim_memo_kernel`, 0.9 s) and with every request walked
    (`sim_kernel`, 3.1 s) -- and must reproduce digest, prefix digests and counters."""
    import hashlib
    path = os.path.join(GOLDEN, "g6_digest.json")
    with open(path) as fh:
        rec = json.load(fh)
    n = int(rec["games"])
    e = _engine_variant(models_s2, stage2="booster", memo=memo)
    try:
        e.set_matchups([MatchupSpec("Kansas State", "Iowa State", KSU, ISU, n, 0, n, 0)])
        got = e.simulate_host(int(rec["seed"]), want_hist=False)
    finally:
        e.close()
    sc = np.ascontiguousarray(got["scores"], dtype=np.int32)
    assert sc.shape == (n, 2)
    for g, digest in rec["sha1_of_first_games"].items():
        assert hashlib.sha1(sc[:int(g)].tobytes()).hexdigest() == digest, f"first {g} games differ from the oracle's"
    assert hashlib.sha1(sc.tobytes()).hexdigest() == rec["sha1_of_int32_scores"]
    for k, v in rec["counters"].items():
        assert got["counters"][k] == v, k
    assert got["counters"]["games"] == n
    assert (got["counters"]["memo_hits"] > 0) == (memo == "on")


@pytest.mark.parametrize("variant", [
    dict(stage2="booster"),
    dict(policy="play_model"),
    dict(sampler="quantile_interp"),
    dict(policy="play_model", stage2="booster", sampler="quantile_interp", play_temp=1.3, qy_noise=0.7),
])
def test_engine_variants_bit_exact(oracle, models_s2, variant):
    """Stage-2 booster (synthetic, trained shape), play_model.xgb policy, sim_helpers sampler."""
    n = 4096
    stream = oracle.make_stream(n, 21)
    e = _engine_variant(models_s2, **variant)
    try:
        e.set_matchups([MatchupSpec("Kansas State", "Iowa State", KSU, ISU, n, 0, n, 0)])
        got = e.simulate_host(0, stream=stream, want_trace=True, want_iters=True)
        cfg = oracle.make_config(
            models_s2, KSU, ISU, policy=variant.get("policy", "heuristic"),
            coach_cols=(e.coach_col("Kansas State"), e.coach_col("Iowa State")),
            play_temp=variant.get("play_temp", 1.0), sampler=variant.get("sampler", "normal"),
            qy_noise=variant.get("qy_noise", 0.5), stage2=variant.get("stage2", "standin"))
        ref = oracle.simulate(cfg, n, stream=stream, trace=True)
        assert np.array_equal(got["scores"], ref["scores"])
        assert _trace_equal(got["trace"], ref["trace"])
        got2 = e.simulate_host(77)
        ref2 = oracle.simulate(cfg, n, seed=77)
        assert np.array_equal(got2["scores"], ref2["scores"])
    finally:
        e.close()


def test_slate_offsets_and_sharding_invariance(engine, oracle, models_s2):
    """Several matchups in one launch; any split of the game-id range gives the same games."""
    pairs = [(KSU, ISU), ((0.0, 28.0, 27.5), (31.7, 41.9, 10.1)), ((-19.3, 17.3, 36.6), (27.9, 40.4, 12.6))]
    games = 3000
    specs, off = [], 0
    for a, b in pairs:
        specs.append(MatchupSpec("a", "b", a, b, games, 0, games, off))
        off += games
    engine.set_matchups(specs)
    whole = engine.simulate_host(5)
    assert whole["hist"].shape[0] == 3 and int(whole["hist"].sum()) == 3 * games
    for m, (a, b) in enumerate(pairs):
        ref = oracle.simulate(oracle.make_config(models_s2, a, b), games, matchup=m, seed=5)
        assert np.array_equal(whole["scores"][m * games:(m + 1) * games], ref["scores"])
        assert np.array_equal(whole["hist"][m], outputs.histogram_from_scores(ref["scores"]))
    # two "ranks": [0,1234) and [1234,3000)
    hist = np.zeros_like(whole["hist"])
    for lo, hi in ((0, 1234), (1234, games)):
        specs, off = [], 0
        for a, b in pairs:
            specs.append(MatchupSpec("a", "b", a, b, games, lo, hi, off))
            off += hi - lo
        engine.set_matchups(specs)
        part = engine.simulate_host(5)
        hist += part["hist"]
        for m in range(3):
            assert np.array_equal(part["scores"][m * (hi - lo):(m + 1) * (hi - lo)],
                                  whole["scores"][m * games + lo:m * games + hi])
    assert np.array_equal(hist, whole["hist"])


def test_edge_sizes(engine):
    for n in (1, 2, 31, 1025):
        engine.set_matchups([MatchupSpec("A", "B", KSU, ISU, n, 0, n, 0)])
        r = engine.simulate_host(3)
        assert r["counters"]["games"] == n and int(r["hist"].sum()) == n
    engine.set_matchups([MatchupSpec("A", "B", KSU, ISU, 0, 5, 5, 0), MatchupSpec("A", "B", ISU, KSU, 7, 0, 7, 0)])
    r = engine.simulate_host(3)
    assert r["counters"]["games"] == 7 and int(r["hist"][0].sum()) == 0 and int(r["hist"][1].sum()) == 7


def test_ks_against_numpy_driven_oracle(engine, oracle, models_s2):
    """Gate G7: the GPU's Philox + u01 + AS241 draws against the oracle fed NumPy's own generator
    (`np.random.default_rng`: PCG64 uniforms and standard normals, FMC:64, 827, 839, 851, 881-889) through the
    injected-stream path -- two-sample KS at alpha = 0.001 on points per team, margin and total, >= 200 k games each.
    A biased uniform mapping or inverse normal on the GPU would separate the two distributions."""
    from scipy.stats import ks_2samp
    n, chunk = 212_992, 16_384                       # 13 chunks of [16384][360][16] float64 draws
    engine.set_matchups([MatchupSpec("A", "B", KSU, ISU, n, 0, n, 0)])
    g = engine.simulate_host(1)["scores"]
    cfg = oracle.make_config(models_s2, KSU, ISU)
    parts = []
    for i in range(n // chunk):
        stream = oracle.make_stream(chunk, 1000 + i)
        parts.append(oracle.simulate(cfg, chunk, game0=i * chunk, stream=stream)["scores"])
        del stream
    o = np.concatenate(parts)
    assert len(o) == n >= 200_000
    for name, a, b in (("ptsA", g[:, 0], o[:, 0]), ("ptsB", g[:, 1], o[:, 1]),
                       ("margin", g[:, 0] - g[:, 1], o[:, 0] - o[:, 1]), ("total", g.sum(1), o.sum(1))):
        assert ks_2samp(a, b).pvalue > 0.001, name


def test_full_size_properties(engine):
    """BASELINE config 2 at full size (10 M games): size-independent invariants."""
    n = 10_000_000
    engine.set_matchups([MatchupSpec("Kansas State", "Iowa State", KSU, ISU, n, 0, n, 0)])
    r = engine.simulate_host(20251018, want_iters=True)
    c = r["counters"]
    assert c["games"] == n and int(r["hist"].sum()) == n and c["hist_overflow"] == 0
    assert c["plays"] == c["pass"] + c["run"]
    assert c["pass"] == c["comp"] + c["inc"] + c["int"] + c["sack"]
    assert c["iters"] == int(r["iters"].sum()) and r["iters"].max() <= 360
    assert c["iters"] == c["plays"] + c["fga"] + c["punt"]
    assert c["fg"] <= c["fga"]
    sc = r["scores"]
    ok = np.zeros(256, dtype=bool)
    for i in range(40):
        for j in range(40):
            if 7 * i + 3 * j < 256:
                ok[7 * i + 3 * j] = True
    assert ok[sc].all()                                   # only touchdowns (7) and field goals (3)
    assert np.array_equal(r["hist"][0], outputs.histogram_from_scores(sc))
    pts = 7 * c["td"] + 3 * c["fg"]
    assert pts == int(sc.sum())
    # both orientations played equally often
    assert int(r["hist"][0, 0].sum()) == n // 2 and int(r["hist"][0, 1].sum()) == n // 2


def test_injected_stream_across_the_slate(engine, oracle, models_s2):
    """Every orientation folds the forests differently (SP+ constants, constant trees, group depths):
    bit-exact trajectories for a spread of real team pairs, strongest vs weakest included."""
    from fast_monte_carlo_b200 import priors
    sp = priors.load_sp_flex(priors.packaged_priors_path()).drop_duplicates(subset=["RATING", "OFFENSE", "DEFENSE"])
    sp = sp.sort_values("RATING").reset_index(drop=True)
    t = [tuple(float(x) for x in sp.loc[i, ["RATING", "OFFENSE", "DEFENSE"]]) for i in range(len(sp))]
    pairs = [(t[-1], t[0]), (t[0], t[-1]), (t[10], t[11]), (t[60], t[100]), (t[-2], t[-3]), (t[33], t[7])]
    n = 1024
    stream = oracle.make_stream(n, 99)
    for a, b in pairs:
        engine.set_matchups([MatchupSpec("A", "B", a, b, n, 0, n, 0)])
        got = engine.simulate_host(0, stream=stream, want_trace=True, want_iters=True)
        ref = oracle.simulate(oracle.make_config(models_s2, a, b), n, stream=stream, trace=True)
        assert np.array_equal(got["scores"], ref["scores"]), (a, b)
        assert np.array_equal(got["iters"], ref["iters"]) and _trace_equal(got["trace"], ref["trace"]), (a, b)


def test_visit_counter_and_gather_probe(engine):
    """FMC_C_VISITS (node slots gathered by live requests) and the roofline probe report sane values."""
    n = 20000
    engine.set_matchups([MatchupSpec("A", "B", KSU, ISU, n, 0, n, 0)])
    c = engine.simulate_host(3)["counters"]
    per_request = c["visits"] / c["requests"]
    assert 100 < per_request < 20000 and c["requests"] + c["memo_hits"] >= c["plays"]
    assert c["warp_steps"] * 32 >= c["visits"] > 0
    l1 = engine.ctx.gather_probe(64 << 10, 500)
    l2 = engine.ctx.gather_probe(8 << 20, 200)
    assert l1 > l2 > 100.0          # GB/s


def test_adversarial_injected_streams(oracle, models_s2):
    """Draw streams that force the rare branches: every uniform at 0 / just below 1, huge normals (sacks that
    push yardsToGoal past 100, negative field position, long fourth-down sequences, down >= 5 chains,
    touchback punts, missed field goals), and per-slot mixtures -- bit-exact against the oracle."""
    from oracle import c_oracle as co
    n = 512
    rng = np.random.default_rng(123)
    base = oracle.make_stream(n, 5)
    normal = list(co.NORMAL_SLOTS)
    uniform = [s for s in range(base.shape[2]) if s not in normal]
    streams = []
    for u, z in ((0.0, -4.0), (1.0 - 2.0 ** -40, 4.0), (0.5, 0.0), (1e-12, 3.0), (0.999, -3.0)):
        s = base.copy()
        s[:, :, uniform] = u
        s[:, :, normal] = z
        streams.append(s)
    for _ in range(3):                      # per-game, per-slot extremes mixed with ordinary draws
        s = base.copy()
        mask = rng.random(s.shape) < 0.35
        ext = np.where(rng.random(s.shape) < 0.5, 0.0, 1.0 - 2.0 ** -33)
        zext = np.where(rng.random(s.shape) < 0.5, -5.0, 5.0)
        s[:, :, uniform] = np.where(mask[:, :, uniform], ext[:, :, uniform], s[:, :, uniform])
        s[:, :, normal] = np.where(mask[:, :, normal], zext[:, :, normal], s[:, :, normal])
        streams.append(s)
    for variant in (dict(stage2="booster"), dict(stage2="booster", policy="play_model", sampler="quantile_interp")):
        e = _engine_variant(models_s2, **variant)
        try:
            e.set_matchups([MatchupSpec("Kansas State", "Iowa State", KSU, ISU, n, 0, n, 0)])
            cfg = oracle.make_config(models_s2, KSU, ISU, policy=variant.get("policy", "heuristic"),
                                     coach_cols=(e.coach_col("Kansas State"), e.coach_col("Iowa State")),
                                     sampler=variant.get("sampler", "normal"), stage2="booster")
            for k, s in enumerate(streams):
                got = e.simulate_host(0, stream=s, want_trace=True, want_iters=True)
                ref = oracle.simulate(cfg, n, stream=s, trace=True)
                assert np.array_equal(got["scores"], ref["scores"]), k
                assert np.array_equal(got["iters"], ref["iters"]) and _trace_equal(got["trace"], ref["trace"]), k
                for c in ("plays", "iters", "pass", "comp", "inc", "int", "sack", "run", "td", "fga", "fg", "punt", "go"):
                    assert got["counters"][c] == ref["counters"][c], (k, c)
        finally:
            e.close()
