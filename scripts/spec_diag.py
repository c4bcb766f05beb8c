"""Debugging tooling: event counters of one small run of a library variant (does it end, does the memo hit?).
    python scripts/spec_diag.py name games [games ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FMC_LIB_PATH"] = os.path.join(ROOT, "build_variants", f"libfmc_{sys.argv[1]}.so")
from fast_monte_carlo_b200 import artifacts as art, synth
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
eng = Engine(synth.with_synthetic_stage2(art.load_default_models()), stage2="booster")
for g in (int(x) for x in sys.argv[2:]):
    eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), g, 0, g, 0)])
    t = time.perf_counter()
    r = eng.simulate_host(20251018, want_hist=False)
    c = r["counters"]
    print(json.dumps({"lib": sys.argv[1], "games": g, "ms": round((time.perf_counter() - t) * 1e3, 1),
                      **{k: c[k] for k in ("plays", "rounds", "requests", "trips", "memo_probes", "memo_hits")}}), flush=True)
