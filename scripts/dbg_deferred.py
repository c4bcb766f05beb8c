"""Debug helper for the deferred-rare-stage experiment: scores of the library under test (memo on) vs its own memo-off kernel."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from fast_monte_carlo_b200 import artifacts as art, synth
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
ms = synth.with_synthetic_stage2(art.load_default_models())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
res = {}
for memo in ("off", "on"):
    eng = Engine(ms, stage2="booster", memo=memo)
    eng.set_matchups([MatchupSpec("A", "B", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), n, 0, n, 0)])
    res[memo] = eng.simulate_host(20251018, want_iters=True, want_trace=(n <= 4000))
    eng.close()
a, b = res["off"], res["on"]
bad = np.flatnonzero((a["scores"] != b["scores"]).any(axis=1) | (a["iters"] != b["iters"]))
print("games", n, "mismatching", len(bad), "hist equal", np.array_equal(a["hist"], b["hist"]),
      "hist==scores(on)", int(b["hist"].sum()), {k: (a["counters"][k], b["counters"][k]) for k in ("games", "plays", "iters", "punt", "fga", "int", "sack") })
if len(bad) and "trace" in a:
    g = int(bad[0])
    ta, tb = a["trace"][g], b["trace"][g]
    k = int(np.flatnonzero(~((ta == tb) | (np.isnan(ta) & np.isnan(tb))).all(axis=1))[0])
    print("first bad game", g, "iters", a["iters"][g], b["iters"][g], "first differing iteration", k)
    for j in range(max(0, k - 2), k + 2):
        print(j, "off", ta[j], "\n   on ", tb[j])
