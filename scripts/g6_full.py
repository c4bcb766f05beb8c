"""Gate G6 as a recorded run: GPU scores == CPU-oracle scores, game by game, on the first N Philox games of
BASELINE configs[1] (Kansas State vs Iowa State, seed 20251018, synthetic stage-2 booster), memo on and off.

    python scripts/g6_full.py [games=1000000] [out.json]

The oracle (oracle/fmc_oracle.c, test infrastructure) runs on the host cores in chunks; prints / writes one JSON record."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_monte_carlo_b200 import artifacts as art, synth
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
from oracle import c_oracle as co

games = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
out = sys.argv[2] if len(sys.argv) > 2 else None
KSU, ISU, SEED = (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), 20251018
ms = synth.with_synthetic_stage2(art.load_default_models())
co.load_models(ms)
cfg = co.make_config(ms, KSU, ISU, stage2="booster")
res = {}
for memo in ("on", "off"):
    eng = Engine(ms, stage2="booster", memo=memo)
    eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", KSU, ISU, games, 0, games, 0)])
    t = time.perf_counter()
    res[memo] = eng.simulate_host(SEED, want_iters=True)
    res[memo]["seconds"] = time.perf_counter() - t
    eng.close()
t = time.perf_counter()
chunk = 100_000
threads = len(os.sched_getaffinity(0))
mism = {"on": 0, "off": 0}
plays = 0
for lo in range(0, games, chunk):
    n = min(chunk, games - lo)
    r = co.simulate(cfg, n, game0=lo, seed=SEED, threads=threads)
    plays += r["counters"]["plays"]
    for memo in ("on", "off"):
        mism[memo] += int((res[memo]["scores"][lo:lo + n] != r["scores"]).any(axis=1).sum())
        mism[memo] += int((res[memo]["iters"][lo:lo + n] != r["iters"]).sum())
cpu_s = time.perf_counter() - t
rec = {"gate": "G6", "games": games, "seed": SEED, "matchup": "Kansas State vs Iowa State, synthetic stage-2 booster",
       "mismatching_games": mism, "gpu_plays": {m: res[m]["counters"]["plays"] for m in res}, "oracle_plays": plays,
       "memo_hit_rate": res["on"]["counters"]["memo_hits"] / max(res["on"]["counters"]["memo_probes"], 1),
       "gpu_seconds_host_call": {m: res[m]["seconds"] for m in res}, "oracle_seconds": cpu_s, "oracle_threads": threads,
       "pass": mism == {"on": 0, "off": 0} and plays == res["on"]["counters"]["plays"] == res["off"]["counters"]["plays"]}
print(json.dumps(rec))
if out:
    json.dump(rec, open(out, "w"), indent=1)
sys.exit(0 if rec["pass"] else 1)
