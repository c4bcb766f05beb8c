"""The binary play-call model path (`play_model.json`, FMC:273-427) on the CPU: the artifact loader, the oracle and
the packer against fixtures recorded from the UNMODIFIED reference running a synthetic `play_model.json`
(tests/golden/make_golden_play_json.py)."""
import json
import os

import numpy as np
import pytest

import packed_walk as pw
from conftest import GOLDEN
from fast_monte_carlo_b200 import artifacts as art, native
from oracle import tree_oracle as to


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLDEN, "play_json.npz"))
    learner = json.loads(bytes(z["model_json"]).decode())["learner"]
    return dict(z=z, learner=learner, features=json.loads(str(z["features"])), classes=json.loads(str(z["classes"])),
                meta=json.loads(str(z["meta"])), pairs=json.loads(str(z["pairs"])), T=float(z["temperature"]))


@pytest.fixture(scope="module")
def models_pj(models_s2, gold):
    forests = dict(models_s2.forests)
    forests["play_binary"] = art.play_binary_from_learner(gold["learner"], gold["features"], gold["classes"], gold["T"])
    art.check_forest(forests["play_binary"])
    return art.ModelSet(forests, source="shipped+synthetic play_model.json")


@pytest.fixture(scope="module")
def oracle_pj(models_pj):
    from oracle import c_oracle as co
    co.build()
    co.load_models(models_pj, play="play_binary")
    yield co
    co.load_models(models_pj, play="play_model")      # leave the slot as the other tests expect it


def test_loader_relays_features_and_folds_categories(models_pj, gold):
    f = models_pj["play_binary"]
    assert f.n_outputs == 2 and f.extra["classes"] == ["pass", "run"] and f.extra["pass_class"] == 0
    assert f.n_features == 18 and f.num_base == 0 and f.n_num == 17 and not f.zero_is_missing
    internal = f.left >= 0
    assert set(np.unique(f.feat[internal])) <= set(range(15)) | {17}          # never half / two_minute
    # categorical nodes: folded on category code 0 (the reference feeds pd.Categorical([coach]) -> code 0)
    trees = gold["learner"]["gradient_booster"]["model"]["trees"]
    n_cat = sum(len(t["categories_nodes"]) for t in trees)
    assert n_cat > 50 and int((f.feat[internal] == 17).sum()) == n_cat
    assert set(np.unique(f.thr[internal & (f.feat == 17)])) == {-1.0, 1.0}
    with pytest.raises(NotImplementedError):
        art._forest_from_xgb_learner("x", gold["learner"], zero_is_missing=False, groups=[], num_base=0, n_num=17)


def test_oracle_pass_probability_equals_reference_wrapper(models_pj, oracle_pj, gold):
    """P(pass) of the reference's own play_call_pass_prob_binary on 400 states."""
    from fast_monte_carlo_b200 import priors
    sp = priors.load_sp_flex(priors.packaged_priors_path())
    st, want = gold["z"]["states"], gold["z"]["p_pass"]
    got = np.zeros(len(st))
    for i, (down, dist, ytg, sd, sec, off, pair) in enumerate(st):
        a, b = gold["pairs"][int(pair)]
        cfg = oracle_pj.make_config(models_pj, priors.lookup_sp_flex(a, sp), priors.lookup_sp_flex(b, sp),
                                    policy="play_json", play_temp=gold["T"])
        got[i] = oracle_pj.play_pass_prob(cfg, int(off), int(down), float(dist), float(ytg), int(sd), int(sec))
    # float32 softmax: NumPy's SIMD float32 exp (what the reference runs, not correctly rounded, build-dependent) and
    # the correctly rounded exp of the oracle / kernel differ in the last bit of P(pass) for about a quarter of the states
    assert np.allclose(got, want, rtol=0, atol=1.2e-7)
    assert (got == want).mean() > 0.5
    assert 0.02 <= want.min() and want.max() <= 0.98 and want.std() > 0.05


def test_oracle_trajectories_with_model_policy(models_pj, oracle_pj, gold):
    z, meta = gold["z"], gold["meta"]
    stream = oracle_pj.make_stream(len(meta), int(z["stream_seed"]))
    for g, m in enumerate(meta):
        cfg = oracle_pj.make_config(models_pj, m["sp_a"], m["sp_b"], policy="play_json", play_temp=gold["T"])
        r = oracle_pj.simulate(cfg, 1, game0=g, stream=stream[g:g + 1], trace=True, threads=1)
        n = int(z["iters"][g])
        assert r["iters"][0] == n, g
        assert np.array_equal(r["trace"][0, :n], z["traces"][g, :n]), g
        f = g & 1
        assert (r["scores"][0, f], r["scores"][0, f ^ 1]) == tuple(z["scores"][g])
    # the model policy matters: the heuristic gives other games on the same draws
    cfg = oracle_pj.make_config(models_pj, meta[0]["sp_a"], meta[0]["sp_b"])
    r = oracle_pj.simulate(cfg, 1, game0=0, stream=stream[0:1], trace=True, threads=1)
    assert not np.array_equal(r["trace"][0, :int(z["iters"][0])], z["traces"][0, :int(z["iters"][0])])


def test_packed_tables_equal_original_trees(models_pj, native_lib):
    from test_pack import _rows
    f = models_pj["play_binary"]
    num = _rows(128, 6)
    fv = np.zeros(17)
    fv[6] = fv[7] = 3.0
    fv[8], fv[9], fv[10], fv[11] = 15.6, 35.7, 20.6, 11.0
    num[:, 6:12] = fv[6:12]
    slots, stream, consts, meta = native.pack_forest_host(f, mode=0, cols=(-1, -1), fold_values=fv)
    got = pw.walk(slots, stream, consts, meta, pw.sim_rows(num, False, None), False, f.base_margin)
    ref = to.raw_margin(f, num, np.full((num.shape[0], 2), -1))
    assert np.array_equal(got, ref)
    slots, stream, consts, meta = native.pack_forest_host(f, mode=1, cols=(-1, -1))
    got = pw.walk(slots, stream, consts, meta, pw.predict_rows(num, False, None), False, f.base_margin)
    assert np.array_equal(got, ref)
