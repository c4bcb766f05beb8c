"""CPU only: the oracle's side of two of the path digests `scripts/spec_experiment.py` prints on the GPU -- "slate" (12
matchups x 40,000 games, seed 11: score table + joint histograms; matchup waves, memo keys that carry the matchup) and
"standin_pm" (100,000 games with the play_model.xgb policy and the stand-in stage 2, seed 5).  Equal digests = identical
games.  ~11 minutes on 8 cores.     python scripts/path_digests_oracle.py"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fast_monte_carlo_b200 import api, artifacts as art, outputs, priors, synth
from fast_monte_carlo_b200.engine import HEAD_COACH_MAP
from oracle import c_oracle as co

co.build()
plain = art.load_default_models()
ms = synth.with_synthetic_stage2(plain)
co.load_models(ms)
t0 = time.time()


def digest(sc, hist):
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(sc, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(hist, dtype=np.uint32).tobytes())
    return h.hexdigest()


sp_df = priors.load_sp_flex(priors.packaged_priors_path())
teams = list(sp_df["team"])[:24]
specs = api.slate_specs([(teams[2 * i], teams[2 * i + 1]) for i in range(12)], 40_000, sp_df)
scores, hists = [], []
for i, m in enumerate(specs):
    r = co.simulate(co.make_config(ms, m.sp_a, m.sp_b, stage2="booster"), 40_000, matchup=i, seed=11)
    scores.append(r["scores"])
    hists.append(outputs.histogram_from_scores(r["scores"]))
out = {"slate": digest(np.concatenate(scores, axis=0), np.stack(hists))}
co.load_models(plain)
g = plain["play_model"].group("coach")
coach = (g.column_of(HEAD_COACH_MAP["Kansas State"]), g.column_of(HEAD_COACH_MAP["Iowa State"]))
r = co.simulate(co.make_config(plain, (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), stage2="standin", policy="play_model",
                               coach_cols=coach), 100_000, seed=5)
out["standin_pm"] = digest(r["scores"], outputs.histogram_from_scores(r["scores"])[None])
out["oracle_seconds"] = round(time.time() - t0)
print(json.dumps(out))
