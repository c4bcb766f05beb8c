import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_monte_carlo_b200 import artifacts as art, synth
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
from oracle import c_oracle as co, tree_oracle as to
ms = synth.with_synthetic_stage2(art.load_default_models())
co.load_models(ms)
KSU = (15.6, 35.7, 20.0); ISU = (11.0, 31.5, 20.6)
n = 4096
stream = co.make_stream(n, 21)
e = Engine(ms, policy="play_model", stage2="standin")
e.set_matchups([MatchupSpec("Kansas State", "Iowa State", KSU, ISU, n, 0, n, 0)])
got = e.simulate_host(0, stream=stream, want_trace=True, want_iters=True)
cfg = co.make_config(ms, KSU, ISU, policy="play_model", coach_cols=(e.coach_col("Kansas State"), e.coach_col("Iowa State")))
ref = co.simulate(cfg, n, stream=stream, trace=True)
tg, tr = got["trace"], ref["trace"]
same = (tg == tr) | (np.isnan(tg) & np.isnan(tr))
bad = np.flatnonzero(~same.all(axis=(1, 2)))
print("bad games", len(bad), bad[:10])
pm = ms["play_model"]
for g in bad[:5]:
    it = int(np.flatnonzero(~same[g].all(axis=1))[0])
    print("game", g, "first diff iter", it)
    print(" prev state gpu", tg[g, it - 1]); print(" prev state ref", tr[g, it - 1])
    print(" next gpu", tg[g, it]); print(" next ref", tr[g, it])
    st = tr[g, it - 1]
    first = g & 1
    off_is_first = st[0] == 1.0
    off = first if off_is_first else first ^ 1
    sp = (KSU, ISU)
    o, d = sp[off], sp[off ^ 1]
    sd = (st[3] - st[4]) if off_is_first else (st[4] - st[3])
    raw = np.array([[st[1], st[5], st[6], float(st[6] <= 20), sd, st[2], 3, 3, o[0], o[1], d[2], d[0]]])
    x = to.play_model_features(pm, raw)
    coach = [e.coach_col("Kansas State"), e.coach_col("Iowa State")][off]
    mo = co.predict("play_model", x, np.array([[coach, -1]]), 5)
    mg = e.predict("play_model", raw, coach=["Chris Klieman", "Matt Campbell"][off])
    print(" margins oracle", mo, "gpu predict", mg, "draw U_call", stream[g, it - 1, 0])
    z = mo[0].astype(np.float32); ez = np.exp(z - z.max()); print(" p_pass", ez[1] / ez.sum())
