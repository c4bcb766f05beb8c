"""The binary play-call model path (`play_model.json`, policy 'play_json') on the GPU, through the C ABI: the
reference's own trajectories with a synthetic model (tests/golden/play_json.npz) and the oracle."""
import numpy as np
import pytest

from fast_monte_carlo_b200.engine import Engine, MatchupSpec
from test_play_json import gold, models_pj, oracle_pj       # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine_pj(models_pj, native_lib):
    e = Engine(models_pj, device=0, policy="play_json", stage2="standin")
    yield e
    e.close()


def test_calibration_temperature_and_pass_class(engine_pj, gold):
    assert engine_pj.play_temp == gold["T"] == 1.3          # calibration.json (FMC:335-337)
    assert engine_pj.coach_col("Kansas State") == -1         # the booster sees category code 0 for every team


def test_reference_golden_trajectories_model_policy(engine_pj, gold):
    from oracle import c_oracle as co
    z, meta = gold["z"], gold["meta"]
    stream = co.make_stream(len(meta), int(z["stream_seed"]))
    for g, m in enumerate(meta):
        engine_pj.set_matchups([MatchupSpec(m["team_a"], m["team_b"], tuple(m["sp_a"]), tuple(m["sp_b"]), 1, g, g + 1, 0)])
        r = engine_pj.simulate_host(0, stream=stream[g:g + 1], want_trace=True, want_iters=True)
        k = int(z["iters"][g])
        assert r["iters"][0] == k, g
        assert np.array_equal(r["trace"][0, :k], z["traces"][g, :k]), g
        f = g & 1
        assert (r["scores"][0, f], r["scores"][0, f ^ 1]) == tuple(z["scores"][g])


def test_margins_bit_exact(engine_pj, oracle_pj, models_pj):
    from test_pack import _rows
    rows = _rows(5000, 11)
    got = engine_pj.predict("play_binary", rows)
    ref = oracle_pj.predict("play_binary", rows, np.full((rows.shape[0], 2), -1), 2)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("stage2", ["standin", "booster"])
def test_injected_stream_vs_oracle(oracle_pj, models_pj, gold, stage2):
    n = 4096
    spa, spb = (15.6, 35.7, 20.0), (11.0, 31.5, 20.6)
    stream = oracle_pj.make_stream(n, 41)
    e = Engine(models_pj, device=0, policy="play_json", stage2=stage2, play_temp=0.9)
    try:
        e.set_matchups([MatchupSpec("Kansas State", "Iowa State", spa, spb, n, 0, n, 0)])
        got = e.simulate_host(0, stream=stream, want_trace=True, want_iters=True)
        cfg = oracle_pj.make_config(models_pj, spa, spb, policy="play_json", play_temp=0.9, stage2=stage2)
        ref = oracle_pj.simulate(cfg, n, stream=stream, trace=True)
        assert np.array_equal(got["scores"], ref["scores"]) and np.array_equal(got["iters"], ref["iters"])
        t0, t1 = got["trace"], ref["trace"]
        assert bool(((t0 == t1) | (np.isnan(t0) & np.isnan(t1))).all())
        for k in ("plays", "pass", "comp", "inc", "int", "sack", "run", "td", "fga", "fg", "punt", "go"):
            assert got["counters"][k] == ref["counters"][k], k
    finally:
        e.close()
