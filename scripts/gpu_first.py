"""First GPU contact: predict parity, injected-trajectory parity, Philox parity, rough timing."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_monte_carlo_b200 import artifacts as art
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
from oracle import c_oracle as co

ms = art.load_default_models()
co.load_models(ms)
eng = Engine(ms)
print("device", eng.ctx.device_name, eng.ctx.sm_count, eng.ctx.smem_per_block, flush=True)
g = np.load("tests/golden/sklearn_quantiles.npz")
num = g["num"].copy(); num[:5, 1] = 0.0; num[5:10, 2] = 0.0
for name in ("pass_stage1", "pass_yards", "run_yards", "sack_yards", "run_fumble", "play_model"):
    f = ms[name]
    rows = num[:, :f.n_num]
    cols = [gg.column_of("Unknown") for gg in f.groups if gg.name != "coach"] + [-1, -1]
    act = np.tile(np.array(cols[:2]), (rows.shape[0], 1))
    x = rows
    if f.scaler_cols is not None:
        from oracle import tree_oracle as to
        x = to.play_model_features(f, rows)
    ref = co.predict(name, x, act, f.n_outputs)
    got = eng.predict(name, rows)
    print(name, "bit-exact", np.array_equal(ref, got), "max abs", np.abs(ref - got).max(), flush=True)

ksu = (15.6, 35.7, 20.0); isu = (11.0, 31.5, 20.6)
# injected trajectories
n = 2048
stream = co.make_stream(n, 7)
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", ksu, isu, n, 0, n, 0)])
t = time.time(); r = eng.simulate_host(0, stream=stream, want_trace=True, want_iters=True); print("gpu injected", time.time() - t)
cfg = co.make_config(ms, ksu, isu)
t = time.time(); o = co.simulate(cfg, n, stream=stream, trace=True); print("oracle injected", time.time() - t)
print("scores equal", np.array_equal(r["scores"], o["scores"]), "iters equal", np.array_equal(r["iters"], o["iters"]))
tr_g = r["trace"]; tr_o = o["trace"]
same = (tr_g == tr_o) | (np.isnan(tr_g) & np.isnan(tr_o))
print("trace bit-exact", bool(same.all()), "mismatching games", int((~same.all(axis=(1, 2))).sum()))
print("counters gpu", r["counters"]); print("counters ora", o["counters"])
# philox
n = 200000
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", ksu, isu, n, 0, n, 0)])
t = time.time(); r = eng.simulate_host(20251018); dt = time.time() - t
print("gpu philox %d games %.3fs -> %.0f games/s %.0f plays/s" % (n, dt, n / dt, r["counters"]["plays"] / dt))
t = time.time(); o = co.simulate(cfg, 20000, seed=20251018); dto = time.time() - t
print("oracle philox 20000 games %.3fs -> %.0f games/s" % (dto, 20000 / dto))
print("philox scores equal (first 20000)", np.array_equal(r["scores"][:20000], o["scores"]),
      int((r["scores"][:20000] != o["scores"]).any(axis=1).sum()))
print("hist sum", int(r["hist"].sum()), "mean pts", r["scores"].mean(axis=0))
for n in (2_000_000,):
    eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", ksu, isu, n, 0, n, 0)])
    t = time.time(); r = eng.simulate_host(1, want_scores=False); dt = time.time() - t
    print("gpu philox %d games %.3fs -> %.0f games/s %.3e plays/s" % (n, dt, n / dt, r["counters"]["plays"] / dt), r["counters"])
print("packed slots", eng.ctx.packed_slots(0).tolist())
