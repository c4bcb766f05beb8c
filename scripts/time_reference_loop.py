"""Times the UNMODIFIED reference module's own simulate_game loop (FMC:1428) in the build container --
SURVEY 8(d)'s "reference-style loop" CPU figure.  Needs /root/reference, so it cannot run on the GPU box;
its output is committed under profiles/ as an offline measurement.  xgboost is not installed: the two
boosters are evaluated by the NumPy stand-in oracle/fake_xgboost.py (the sklearn quantile pipelines, which
dominate the cost, are the reference's own objects).

    python scripts/time_reference_loop.py [games=24]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_harness as rh

games = int(sys.argv[1]) if len(sys.argv) > 1 else 24
out = {}
for label, cache_off in (("memo_caches_on_as_shipped", False), ("memo_caches_off", True)):
    rh._MOD = None
    mod = rh.load_reference(cache_off=cache_off)
    mod.RNG = np.random.default_rng(20251018)
    a, b = rh.team_context(mod, "Kansas State"), rh.team_context(mod, "Iowa State")
    mod.simulate_game(a, b, seed=None)          # warm-up (first Pipeline.predict is slow)
    plays = 0
    t = time.perf_counter()
    for g in range(games):
        r = mod.simulate_game(a, b, seed=None) if g % 2 == 0 else mod.simulate_game(b, a, seed=None)
        plays += sum(v["plays"] for v in r["box"].values())
    dt = time.perf_counter() - t
    out[label] = dict(games=games, seconds=dt, games_per_sec=games / dt, plays_per_sec=plays / dt,
                      plays_per_game=plays / games)
print(json.dumps({"what": "reference fast_monte_carlo_cfb.simulate_game, single process, this container "
                          f"({os.cpu_count()} cpus), NumPy stand-in for xgboost", **out}))
