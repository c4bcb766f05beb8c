"""The oracle against everything the reference can tell us (CPU only).

Golden files come from the real reference objects / module (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import tree_oracle as to


def test_philox_known_answers(oracle):
    # Random123 known-answer vectors for philox4x32-10
    assert oracle.philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert oracle.philox((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert oracle.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_inverse_normal_accuracy(oracle):
    from scipy.stats import norm
    ps = np.concatenate([np.linspace(1e-10, 1 - 1e-10, 20001), [0.5, 0.075, 0.925, 2.0 ** -33]])
    got = np.array([oracle.lib().fo_ppnd16(float(p)) for p in ps])
    assert np.abs(got - norm.ppf(ps)).max() < 1e-13


def test_sklearn_quantiles_bitwise(models, oracle):
    """NumPy and C restatements of GradientBoostingRegressor.predict == the live pipelines, bit for bit."""
    g = np.load(os.path.join(GOLDEN, "sklearn_quantiles.npz"))
    num = g["num"]
    for fam in ("pass_yards", "run_yards", "sack_yards"):
        act, pred = g[f"{fam}/active"], g[f"{fam}/pred"]
        assert np.array_equal(to.raw_margin(models[fam], num[:500], act[:500]), pred[:500])
        assert np.array_equal(oracle.predict(fam, num, act, 3), pred)


def test_xgb_provisional_vectors(models, oracle):
    """SURVEY Appendix G vectors (xgboost itself is unavailable: parity unpinned, see oracle header)."""
    p = json.load(open(os.path.join(GOLDEN, "xgb_provisional.json")))
    f = models["pass_stage1"]
    for v in p["stage1"]:
        num = np.array([p["r0"]], dtype=float)
        num[0, 4] = v["score_diff"]
        act = np.array([[f.groups[0].column_of(v["passer"]), -1]])
        m = oracle.predict("pass_stage1", num, act, 1, 0, v["trees"])[0, 0]
        assert abs(m - v["margin"]) < 2e-7
        assert abs(float(to.sigmoid_f32(np.array([m]))[0]) - v["p"]) < 2e-7
        assert np.array_equal(to.raw_margin(f, num, act, 0, v["trees"]), [[m]])
    pm = models["play_model"]
    for v in p["play_model"]:
        raw = np.array([[v["down"], v["distance"], v["ytg"], 0, v["sd"], v["sec"], 3, 3, 15.6, 35.7, 20.6, 11.0]], dtype=float)
        x = to.play_model_features(pm, raw)
        m = oracle.predict("play_model", x, np.array([[-1, -1]]), 5)
        assert np.abs(m[0] - np.array(v["margins"])).max() < 2e-3
    for fam, q in p["sklearn_r0_unknown"].items():
        f = models[fam]
        num = np.array([p["r0"]], dtype=float)
        num[0, 4] = 0
        cols = [g.column_of("Unknown") for g in f.groups] + [-1]
        m = oracle.predict(fam, num, np.array([cols[:2]]), 3)
        assert np.abs(m[0] - np.array(q)).max() < 1e-7


def test_scalar_helpers_match_reference(oracle):
    """pass_prob_v1 / go_for_it_prob / field_goal_prob / modifiers evaluated by the reference module."""
    sc = json.load(open(os.path.join(GOLDEN, "ref_scalars.json")))
    L = oracle.lib()
    for down, dist, ytg, sec, sd, want in sc["pass_prob_v1"]:
        assert L.fo_pass_prob_v1(int(down), dist, ytg, int(sec), int(sd)) == want
    for ytg, dist, sd, sec, want in sc["go_for_it_prob"]:
        assert L.fo_go_for_it_prob(ytg, dist, int(sd), int(sec)) == want
    for ytg, want in sc["field_goal_prob"]:
        assert L.fo_field_goal_prob(ytg + 17) == want
    for m in sc["modifiers"]:
        got = oracle.modifiers(m["off"][1], m["de"][2], m["ytg"], m["down"])
        for k, v in got.items():
            assert v == m[k], (k, v, m[k])


def test_reference_trajectories_bit_exact(models, oracle):
    """Per-iteration states of the reference's own simulate_game under injected draws."""
    t = np.load(os.path.join(GOLDEN, "ref_trajectories.npz"))
    meta = json.loads(str(t["meta"]))
    stream = oracle.make_stream(len(meta), int(t["stream_seed"]))
    for g, m in enumerate(meta):
        spA, spB = (m["sp_second"], m["sp_first"]) if g & 1 else (m["sp_first"], m["sp_second"])
        cfg = oracle.make_config(models, spA, spB)
        r = oracle.simulate(cfg, 1, game0=g, stream=stream[g:g + 1], trace=True, threads=1)
        n = int(t["iters"][g])
        assert r["iters"][0] == n
        assert np.array_equal(r["trace"][0, :n], t["traces"][g, :n])       # ints AND float bits
        first = g & 1
        assert (r["scores"][0, first], r["scores"][0, first ^ 1]) == tuple(t["scores"][g])
        assert r["counters"]["plays"] == int(t["plays"][g])


def test_reference_trajectories_adversarial_streams(models, oracle):
    """The reference's own simulate_game under ADVERSARIAL draws (tests/streams.py): uniforms pinned at 0 / just
    below 1, +-4 sigma normals, mixtures -- the rare branches (touchback punts, missed field goals, sacks past
    the 100-yard line, down >= 5 chains, 112-112 shoot-outs), bit for bit."""
    from streams import adversarial_stream
    t = np.load(os.path.join(GOLDEN, "ref_trajectories_adversarial.npz"))
    meta = json.loads(str(t["meta"]))
    stream = adversarial_stream(len(meta), int(t["stream_seed"]))
    assert int(t["scores"].max()) >= 112 and int(t["scores"].min()) == 0
    for g, m in enumerate(meta):
        spA, spB = (m["sp_second"], m["sp_first"]) if g & 1 else (m["sp_first"], m["sp_second"])
        cfg = oracle.make_config(models, spA, spB)
        r = oracle.simulate(cfg, 1, game0=g, stream=stream[g:g + 1], trace=True, threads=1)
        n = int(t["iters"][g])
        assert r["iters"][0] == n, (g, m["pattern"])
        assert np.array_equal(r["trace"][0, :n], t["traces"][g, :n]), (g, m["pattern"])
        first = g & 1
        assert (r["scores"][0, first], r["scores"][0, first ^ 1]) == tuple(t["scores"][g])
        assert r["counters"]["plays"] == int(t["plays"][g])


def test_oracle_threads_and_determinism(models, oracle):
    cfg = oracle.make_config(models, (15.6, 35.7, 20.0), (11.0, 31.5, 20.6))
    a = oracle.simulate(cfg, 64, seed=5, threads=1)
    b = oracle.simulate(cfg, 64, seed=5, threads=4)
    c = oracle.simulate(cfg, 32, game0=32, seed=5)
    assert np.array_equal(a["scores"], b["scores"])
    assert np.array_equal(a["scores"][32:], c["scores"])
    assert not np.array_equal(a["scores"], oracle.simulate(cfg, 64, seed=6)["scores"])
    # scores are sums of 7s and 3s
    ok = {7 * i + 3 * j for i in range(20) for j in range(20)}
    assert set(a["scores"].ravel().tolist()) <= ok


def test_g6_digest_fixture_is_the_oracles(models_s2, oracle):
    """tests/golden/g6_digest.json (scripts/g6_digest.py: the oracle's 10,000,000 Philox games of BASELINE configs[1],
    2.6 hours of host time, recorded once) is what the GPU test `test_g6_full_size_digest` compares against.  Here the
    oracle replays the first 20,000 of those games: their digest must be the recorded prefix digest, and the record must
    be complete and agree with the GPU run it quotes."""
    import hashlib
    from conftest import ISU, KSU
    with open(os.path.join(GOLDEN, "g6_digest.json")) as fh:
        rec = json.load(fh)
    assert rec["games"] == 10_000_000 and rec["seed"] == 20251018
    assert rec["sha1_of_first_games"][str(rec["games"])] == rec["sha1_of_int32_scores"]
    assert rec["gpu_run"]["sha1_of_int32_scores"] == rec["sha1_of_int32_scores"]
    assert rec["gpu_run"]["plays"] == rec["counters"]["plays"]
    n = 20_000
    r = oracle.simulate(oracle.make_config(models_s2, KSU, ISU, stage2="booster"), n, seed=rec["seed"])
    sc = np.ascontiguousarray(r["scores"], dtype=np.int32)
    assert hashlib.sha1(sc.tobytes()).hexdigest() == rec["sha1_of_first_games"][str(n)]
