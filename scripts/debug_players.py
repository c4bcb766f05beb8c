"""Debug helper: player-mode engine vs oracle on Philox draws; prints where the player boxes differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_monte_carlo_b200 import artifacts as art, synth, priors, usage
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
from oracle import c_oracle as co
ms = synth.with_synthetic_stage2(art.load_default_models())
co.build(); co.load_models(ms)
gold = os.path.join(ROOT, "tests", "golden")
focus = usage.build_focus_usage_tables(os.path.join(gold, "players_focus.csv"))
sp = priors.load_sp_flex(priors.packaged_priors_path())
tcs = [priors.build_team_context_from_sp_flex(t, 2025, 1, sp, focus=focus, usage_dir=gold) for t in ("Kansas State", "Iowa State")]
us = tuple(usage.resolve_team(tc, ms) for tc in tcs)
for stage2 in ("standin", "booster"):
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    e = Engine(ms, device=0, stage2=stage2)
    e.set_matchups([MatchupSpec("Kansas State", "Iowa State", tcs[0].sp, tcs[1].sp, n, 0, n, 0, usage=us)])
    got = e.simulate_host(99, want_iters=True, want_players=True)
    cfg = co.make_config(ms, tcs[0].sp, tcs[1].sp, stage2=stage2)
    ref = co.simulate(cfg, n, seed=99, usage=co.make_usage(us), n_slots=e.n_slots)
    print(stage2, "scores equal", np.array_equal(got["scores"], ref["scores"]), "players equal", np.array_equal(got["players"], ref["players"]))
    d = np.argwhere(got["players"] != ref["players"])
    print("differing entries", len(d), "games", len(set(d[:, 0].tolist())))
    for row in d[:12]:
        g, t, s, f = row
        print(" game", g, "team", t, "slot", s, us[t].slots[s] if s < len(us[t].slots) else None, "field", f,
              "gpu", repr(got["players"][g, t, s, f]), "ref", repr(ref["players"][g, t, s, f]), "iters", got["iters"][g])
    if len(d):
        print(" fields histogram", np.bincount(d[:, 3], minlength=6).tolist(), "slots", np.bincount(d[:, 2], minlength=8).tolist())
    e.close()
