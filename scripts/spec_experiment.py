"""Measurement tooling: one build of libfmc_b200.so (scripts/build_variants.py) on the headline workload, without torch
(a fresh box pays most of a minute for `import torch`):   python scripts/spec_experiment.py games name
1 warm + 3 timed `fmc_simulate_host` calls of `games` games of configs[1] (host wall clock: kernel + the copy of the
per-game score words), a SHA-1 of the score table (every build must print the same one), then digests of four other code
paths of the memo kernel.  One JSON line per measurement, with the seconds since start."""
import hashlib, json, os, sys, time
T0 = time.perf_counter()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
games, name = int(sys.argv[1]), sys.argv[2]
os.environ["FMC_LIB_PATH"] = os.path.join(ROOT, "build_variants", f"libfmc_{name}.so")
from fast_monte_carlo_b200 import artifacts as art, synth
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
t_import = time.perf_counter() - T0
ms = synth.with_synthetic_stage2(art.load_default_models())
t_models = time.perf_counter() - T0
eng = Engine(ms, stage2="booster")
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), games, 0, games, 0)])


def timed(n):
    out, res = [], None
    for _ in range(n):
        t = time.perf_counter()
        res = eng.simulate_host(20251018, want_hist=False)
        out.append(round((time.perf_counter() - t) * 1e3, 1))
    return out, res


w, _ = timed(1)
t, res = timed(3)
c = res["counters"]
print(json.dumps({"lib": name, "games": games, "warm_ms": w, "ms": t, "plays": int(c["plays"]), "plays_per_s_e2e": int(c["plays"]) / min(t) * 1e3,
                  "memo_hits": int(c["memo_hits"]), "memo_probes": int(c["memo_probes"]),
                  "sha1": hashlib.sha1(res["scores"].tobytes()).hexdigest(), "t_import_s": round(t_import, 1),
                  "t_models_s": round(t_models, 1), "t_done_s": round(time.perf_counter() - T0, 1)}), flush=True)
# regression digests over the other code paths of the memo kernel (every build must print the same ones): the parity-test
# instantiation (injected draws, per-iteration trace), a slate (matchup waves), stand-in stage 2 + the play-model policy
import numpy as np
from fast_monte_carlo_b200 import api, priors
from fast_monte_carlo_b200.native import MAX_ITERS, N_SLOTS


def sha(*arrs):
    h = hashlib.sha1()
    for a_ in arrs:
        h.update(np.ascontiguousarray(a_).tobytes())
    return h.hexdigest()


def stream_of(n, seed):
    rng = np.random.default_rng(seed)
    s_ = rng.random((n, MAX_ITERS, N_SLOTS))
    for sl in (2, 7, 10, 11):
        s_[:, :, sl] = rng.standard_normal((n, MAX_ITERS))
    return s_


dig = {}
n_inj = 512
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), n_inj, 0, n_inj, 0)])
r_ = eng.simulate_host(0, stream=stream_of(n_inj, 7), want_trace=True, want_iters=True)
dig["injected"] = sha(r_["scores"], r_["iters"], r_["trace"])
sp_df = priors.load_sp_flex(priors.packaged_priors_path())
teams = list(sp_df["team"])[:24]
eng.set_matchups(api.slate_specs([(teams[2 * i], teams[2 * i + 1]) for i in range(12)], 40_000, sp_df))
r_ = eng.simulate_host(11)
dig["slate"] = sha(r_["scores"], r_["hist"])
eng.close()
eng = Engine(art.load_default_models(), stage2="standin", policy="play_model")
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), 100_000, 0, 100_000, 0)])
r_ = eng.simulate_host(5)
dig["standin_pm"] = sha(r_["scores"], r_["hist"])
dig["standin_pm_hits"] = int(r_["counters"]["memo_hits"])
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), 128, 0, 128, 0)])
r_ = eng.simulate_host(0, stream=stream_of(128, 9), want_trace=True, want_iters=True)
dig["standin_pm_injected"] = sha(r_["scores"], r_["iters"], r_["trace"])
print(json.dumps({"lib": name, "digests": dig, "t_done_s": round(time.perf_counter() - T0, 1)}), flush=True)
eng.close()
