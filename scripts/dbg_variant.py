import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from fast_monte_carlo_b200 import artifacts as art, synth
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
ms = synth.with_synthetic_stage2(art.load_default_models())
for n in (32, 1000, 100000):
    eng = Engine(ms, stage2="booster")
    eng.set_matchups([MatchupSpec("A", "B", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), n, 0, n, 0)])
    try:
        r = eng.simulate_host(1)
        print(n, "ok", r["counters"]["games"], r["counters"]["memo_hits"], r["counters"]["requests"], flush=True)
    except Exception as e:
        print(n, "FAIL", str(e)[:300], flush=True); break
