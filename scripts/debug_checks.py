"""Runs simulations and predict calls on a library built with -DFMC_DEBUG_CHECKS (every node gather and feature
offset bounds-checked on the device) and reports fmc_debug_errors():
    nvcc ... -DFMC_DEBUG_CHECKS -o build_variants/libfmc_dbg.so fast_monte_carlo_b200/csrc/fmc_abi.cu
    FMC_LIB_PATH=$PWD/build_variants/libfmc_dbg.so python scripts/debug_checks.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_monte_carlo_b200 import artifacts as art, synth, native
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
ms = synth.with_synthetic_stage2(art.load_default_models())
L = native.load_library()
total = 0
for kw in (dict(stage2="booster"), dict(stage2="standin"), dict(stage2="booster", policy="play_model", sampler="quantile_interp")):
    eng = Engine(ms, **kw)
    specs = [MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), 40000, 0, 40000, 0),
             MatchupSpec("UTSA", "Ohio State", (0.0, 28.0, 27.5), (31.7, 41.9, 10.1), 15000, 0, 15000, 40000),
             MatchupSpec("Best", "Worst", (31.7, 41.9, 10.1), (-19.3, 17.3, 36.6), 7, 0, 7, 55000)]
    eng.set_matchups(specs)
    r = eng.simulate_host(5, want_iters=True)
    e = int(L.fmc_debug_errors()); total += e
    print(kw, "games", r["counters"]["games"], "plays", r["counters"]["plays"], "debug errors", e)
    rng = np.random.default_rng(0)
    rows = rng.normal(5, 10, (5000, 17)); rows[:, 0] = rng.integers(1, 7, 5000); rows[::3, 4] = 0.0; rows[::5, 1] = 0.0
    for name in ("pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards", "run_fumble", "play_model"):
        eng.predict(name, rows[:, :ms[name].n_num])
        e = int(L.fmc_debug_errors()); total += e
        print("  predict", name, "debug errors", e)
    eng.close()
big = synth.synthetic_stage2(ms, seed=5, rounds=1500)
forests = dict(ms.forests); forests["pass_stage2"] = big
eng = Engine(art.ModelSet(forests, source="big"), stage2="standin")
eng.predict("pass_stage2", rows)
e = int(L.fmc_debug_errors()); total += e
print("multi-window forest: debug errors", e)
# player mode (dynamic one-hot rows, running box, per-player histograms) and the play_model.json policy
from fast_monte_carlo_b200 import priors, usage
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
focus = usage.build_focus_usage_tables(os.path.join(gold, "players_focus.csv"))
sp = priors.load_sp_flex(priors.packaged_priors_path())
tcs = {t: priors.build_team_context_from_sp_flex(t, 2025, 1, sp, focus=focus, usage_dir=gold)
       for t in ("Kansas State", "Iowa State", "Ohio State", "UTSA")}
learner = art.play_binary_from_learner(synth.synthetic_play_model_json()["learner"], synth.PLAY_JSON_FEATURES, ["pass", "run"], 1.3)
forests = dict(ms.forests); forests["play_binary"] = learner
msp = art.ModelSet(forests, source="players")
for kw in (dict(stage2="booster"), dict(stage2="booster", policy="play_json")):
    eng = Engine(msp, **kw)
    specs = []
    off = 0
    for a, b, n in (("Kansas State", "Iowa State", 30000), ("Ohio State", "UTSA", 9000), ("Ohio State", "Kansas State", 5)):
        specs.append(MatchupSpec(a, b, tcs[a].sp, tcs[b].sp, n, 0, n, off,
                                 usage=(usage.resolve_team(tcs[a], msp), usage.resolve_team(tcs[b], msp))))
        off += n
    eng.set_matchups(specs)
    r = eng.simulate_host(9, want_players=True, want_player_hist=True)
    e = int(L.fmc_debug_errors()); total += e
    print("players", kw, "games", r["counters"]["games"], "box lines", int(r["players"].dense()[..., 1].sum()),
          "hist", int(r["player_hist"].sum()), "debug errors", e)
    eng.close()
print("TOTAL debug errors", total)
sys.exit(1 if total else 0)
