"""Gate G6 at BASELINE configs[1]'s full size without a GPU-side oracle run: the C oracle plays all N Philox games of the
headline workload (Kansas State - Iowa State, synthetic stage-2 booster, seed 20251018) on the host cores, in chunks, and
records the SHA-1 of the per-game score table (int32 [N, 2], the layout `Context.simulate_host` returns) plus its event
counters.  The GPU side is one `simulate_host` call of the same N games (0.9 s): `tests/test_gpu_sim.py` compares its
digest with the recorded one.     python scripts/g6_digest.py [games] [out.json] [chunk]"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fast_monte_carlo_b200 import artifacts as art, synth
from oracle import c_oracle as co

games = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", "g6_digest.json")
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 250_000
SEED = 20251018
co.build()
ms = synth.with_synthetic_stage2(art.load_default_models())
co.load_models(ms)
cfg = co.make_config(ms, (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), stage2="booster")
h = hashlib.sha1()
tot = {}
marks = {}
t0 = time.time()
g = 0
while g < games:
    n = min(chunk, games - g)
    r = co.simulate(cfg, n, game0=g, seed=SEED)
    h.update(np.ascontiguousarray(r["scores"], dtype=np.int32).tobytes())
    for k, v in r["counters"].items():
        tot[k] = tot.get(k, 0) + int(v)
    g += n
    if g in (100_000, 1_000_000) or g == games:
        marks[str(g)] = h.copy().hexdigest()      # digests of the first 100 k / 1 M games: cheap gates for small runs
    print(f"{g} games, {time.time() - t0:.0f} s", flush=True)
rec = {"workload": "configs[1]: Kansas State vs Iowa State, synthetic stage-2 booster, heuristic play call, Philox seed 20251018",
       "seed": SEED, "games": games, "sha1_of_int32_scores": h.hexdigest(), "sha1_of_first_games": marks, "counters": tot,
       "oracle_seconds": round(time.time() - t0, 1), "host_threads": os.cpu_count(),
       "how": "oracle/fmc_oracle.c (fo_simulate, OpenMP), chunks of %d games hashed in game order" % chunk}
with open(out, "w") as f:
    json.dump(rec, f, indent=1)
print(json.dumps(rec))
