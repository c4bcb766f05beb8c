"""Synthetic stand-ins for artifacts the reference snapshot does not ship.

`pass_stage2_notcomplete.json` is listed in the reference's .MISSING_LARGE_BLOBS, yet the engine
loads it at import (FMC:642) and evaluates it on every incomplete pass (FMC:751-770).  To exercise
and time that path, `synthetic_stage2` builds a booster of the trained SHAPE
(train_pass_outcome_stage2.py:42-43, 94-106; pass_stage2_meta.json best_iteration=241):
multi:softprob, 3 classes [incomplete, intercepted, sack], 362 rounds = 1,086 trees of depth <= 7
over the 504-column stage-2 layout, whose split structure is resampled from the shipped stage-1
booster and whose margins hover around the class priors.  It is NOT the trained model; results
obtained with it say nothing about football, only about throughput and parity.
"""
from __future__ import annotations

import numpy as np

from . import artifacts as art

STAGE2_ROUNDS = 362          # best_iteration 241 + 1 + early-stopping patience 120
STAGE2_PRIORS = (0.78, 0.05, 0.17)


def synthetic_stage2(ms: art.ModelSet, seed: int = 2025, rounds: int = STAGE2_ROUNDS,
                     groups=None) -> art.Forest:
    s1 = ms["pass_stage1"]
    rng = np.random.default_rng(seed)
    if groups is None:
        # stage-2 layout of the shipped preprocessor: 486 passers ("Unknown" at 467), 1 target, 17 numerics
        passers = [f"passer_{i:03d}" for i in range(486)]
        passers[467] = "Unknown"
        groups = [art.OneHotGroup("passer_name", 0, passers), art.OneHotGroup("target_name", 486, ["Unknown"])]
    n_onehot = sum(len(g.categories) for g in groups)
    num_base = n_onehot
    n_features = num_base + art.N_NUM
    n_pass = len(groups[0].categories)
    feat, thr, left, right, dl, val, cov, roots, outs = [], [], [], [], [], [], [], [], []
    off = 0
    n_s1 = s1.n_trees
    bounds = np.append(s1.tree_root, s1.n_nodes)
    for r in range(rounds):
        for k in range(3):
            t = int((3 * r + k + rng.integers(0, n_s1)) % n_s1)
            a, b = int(bounds[t]), int(bounds[t + 1])
            f = s1.feat[a:b].astype(np.int64)
            leaf = s1.left[a:b] < 0
            is_num = f >= s1.num_base
            f2 = np.where(is_num, f - s1.num_base + num_base, f % n_pass)
            roots.append(off)
            outs.append(k)
            feat.append(np.where(leaf, -1, f2).astype(np.int32))
            thr.append(s1.thr[a:b].copy())
            left.append(np.where(leaf, -1, s1.left[a:b] - a + off).astype(np.int32))
            right.append(np.where(leaf, -1, s1.right[a:b] - a + off).astype(np.int32))
            dl.append(s1.default_left[a:b].copy())
            v = s1.value[a:b] * 0.2 * rng.choice([-1.0, 1.0]) * rng.uniform(0.5, 1.0)
            if r == 0:
                v = v * 0.0 + (np.log(STAGE2_PRIORS[k]) - 0.5)
            val.append(np.where(leaf, v.astype(np.float32).astype(np.float64), 0.0))
            cov.append(s1.cover[a:b].copy() if s1.cover is not None else np.zeros(b - a))
            off += b - a
    fo = art.Forest(
        name="pass_stage2", kind=art.KIND_XGB, link=art.LINK_SOFTMAX, n_outputs=3, n_features=n_features,
        num_base=num_base, n_num=art.N_NUM, zero_is_missing=True,
        base_margin=np.full(3, 0.5, dtype=np.float64), scale=1.0, groups=groups,
        feat=np.concatenate(feat), thr=np.concatenate(thr), left=np.concatenate(left),
        right=np.concatenate(right), default_left=np.concatenate(dl), value=np.concatenate(val),
        tree_root=np.asarray(roots, dtype=np.int32), tree_out=np.asarray(outs, dtype=np.int32),
        cover=np.concatenate(cov), best_iteration=241)
    fo.extra["synthetic"] = True
    art.check_forest(fo)
    return fo


def with_synthetic_stage2(ms: art.ModelSet, seed: int = 2025) -> art.ModelSet:
    forests = dict(ms.forests)
    forests["pass_stage2"] = synthetic_stage2(ms, seed)
    return art.ModelSet(forests, source=ms.source + "+synthetic_stage2")


def xgb_json_dict(f: art.Forest) -> dict:
    """A Forest in the XGBoost JSON model schema (subset the loader reads) -- lets tests push a
    synthetic booster through the same parser as the shipped ones."""
    trees = []
    bounds = np.append(f.tree_root, f.n_nodes)
    for t in range(f.n_trees):
        a, b = int(bounds[t]), int(bounds[t + 1])
        leaf = f.left[a:b] < 0
        sc = np.where(leaf, f.value[a:b], f.thr[a:b].astype(np.float64))
        trees.append(dict(
            left_children=[int(x) for x in np.where(leaf, -1, f.left[a:b] - a)],
            right_children=[int(x) for x in np.where(leaf, -1, f.right[a:b] - a)],
            split_indices=[int(x) for x in np.where(leaf, 0, f.feat[a:b])],
            split_conditions=[float(np.float32(x)) for x in sc],
            default_left=[int(x) for x in f.default_left[a:b]],
            split_type=[0] * (b - a),
            sum_hessian=[float(x) for x in (f.cover[a:b] if f.cover is not None else np.zeros(b - a))],
            base_weights=[float(np.float32(x)) for x in sc],
        ))
    objective = {art.LINK_SIGMOID: "binary:logistic", art.LINK_SOFTMAX: "multi:softprob"}.get(f.link, "reg:squarederror")
    base = float(f.base_margin[0]) if f.link != art.LINK_SIGMOID else float(1.0 / (1.0 + np.exp(-f.base_margin[0])))
    return dict(learner=dict(
        attributes=dict(best_iteration=str(f.best_iteration)) if f.best_iteration is not None else {},
        feature_names=[], feature_types=[],
        gradient_booster=dict(name="gbtree", model=dict(
            gbtree_model_param=dict(num_parallel_tree="1", num_trees=str(f.n_trees)),
            iteration_indptr=list(range(0, f.n_trees + 1, max(1, f.n_outputs))),
            tree_info=[int(x) for x in f.tree_out], trees=trees)),
        learner_model_param=dict(base_score=repr(base), boost_from_average="1",
                                 num_class=str(f.n_outputs if f.n_outputs > 1 else 0),
                                 num_feature=str(f.n_features), num_target="1"),
        objective=dict(name=objective)), version=[3, 0, 4])


# ----------------------------------------------------------------------------------------------
# play_model.json (the binary PASS/RUN booster of train_run_pass.py) is not in the reference snapshot either.
# ----------------------------------------------------------------------------------------------
PLAY_JSON_FEATURES = ["down", "distance", "yardsToGoal", "is_red_zone", "score_diff", "seconds_remaining",
                      "offenseTimeouts", "defenseTimeouts", "sp_rating_off", "sp_offense_rating_off",
                      "sp_defense_rating_def", "sp_rating_def", "head_coach", "goal_to_go", "fourth_and_short",
                      "fg_range"]                      # features.pkl


def synthetic_play_model_json(seed: int = 7, rounds: int = 60, max_depth: int = 6) -> dict:
    """A booster of the trained SHAPE of play_model.json (train_run_pass.py:171-199: multi:softprob, 2 classes
    [pass, run], max_depth 6, 16 features with the categorical `head_coach`), in the XGBoost JSON schema, with
    random split structure: numeric splits on the 15 numerics at plausible raw thresholds and categorical
    splits on `head_coach` (split_type 1, category sets that do / do not contain code 0).  NOT a trained model."""
    rng = np.random.default_rng(seed)
    n_feat = len(PLAY_JSON_FEATURES)

    def threshold(c):
        nm = PLAY_JSON_FEATURES[c]
        if nm == "down":
            return float(rng.choice([1.5, 2.5, 3.5]))
        if nm == "distance":
            return float(np.round(rng.uniform(0.5, 15.0), 2))
        if nm == "yardsToGoal":
            return float(np.round(rng.uniform(1.0, 99.0), 1))
        if nm == "score_diff":
            return float(rng.integers(-21, 22)) + float(rng.choice([0.0, 0.5]))   # incl. thresholds exactly at 0
        if nm == "seconds_remaining":
            return float(rng.integers(30, 3600))
        if nm in ("offenseTimeouts", "defenseTimeouts"):
            return float(rng.choice([1.5, 2.5, 3.0]))
        if nm.startswith("sp_"):
            return float(np.round(rng.uniform(-25.0, 45.0), 1))
        return 0.5                                      # 0/1 flags

    trees, info = [], []
    for r in range(rounds):
        for k in range(2):
            lc, rc, si, sc, dl, st = [], [], [], [], [], []
            cats, cnodes, cseg, csz = [], [], [], []
            depth_of = [0]
            lc.append(-1); rc.append(-1); si.append(0); sc.append(0.0); dl.append(0); st.append(0)
            i = 0
            while i < len(lc):
                d = depth_of[i]
                if d < max_depth and rng.random() < (0.95 if d < 2 else 0.62):
                    c = int(rng.integers(0, n_feat))
                    a = len(lc)
                    for _ in range(2):
                        lc.append(-1); rc.append(-1); si.append(0); sc.append(0.0); dl.append(0); st.append(0)
                        depth_of.append(d + 1)
                    lc[i], rc[i], si[i], dl[i] = a, a + 1, c, int(rng.integers(0, 2))
                    if PLAY_JSON_FEATURES[c] == "head_coach":
                        st[i] = 1
                        members = sorted(set(int(x) for x in rng.integers(0, 6, size=int(rng.integers(1, 4)))))
                        cnodes.append(i); cseg.append(len(cats)); csz.append(len(members)); cats.extend(members)
                        sc[i] = float(len(members))
                    else:
                        sc[i] = threshold(c)
                else:
                    sc[i] = float(np.float32(rng.normal(0.0, 0.15) + (0.1 if k == 0 else -0.1) * (r == 0)))
                i += 1
            trees.append(dict(left_children=lc, right_children=rc, split_indices=si,
                              split_conditions=[float(np.float32(x)) for x in sc], default_left=dl, split_type=st,
                              categories=cats, categories_nodes=cnodes, categories_segments=cseg, categories_sizes=csz,
                              sum_hessian=[1.0] * len(lc), base_weights=[float(np.float32(x)) for x in sc],
                              tree_param=dict(num_nodes=str(len(lc)), num_feature=str(n_feat))))
            info.append(k)
    return dict(learner=dict(
        attributes=dict(best_iteration=str(rounds - 1)),
        feature_names=list(PLAY_JSON_FEATURES),
        feature_types=["c" if n == "head_coach" else ("float" if n in ("distance", "yardsToGoal") or n.startswith("sp_") else "int")
                       for n in PLAY_JSON_FEATURES],
        gradient_booster=dict(name="gbtree", model=dict(
            gbtree_model_param=dict(num_parallel_tree="1", num_trees=str(len(trees))),
            iteration_indptr=list(range(0, len(trees) + 1, 2)), tree_info=info, trees=trees)),
        learner_model_param=dict(base_score="5E-1", boost_from_average="1", num_class="2",
                                 num_feature=str(n_feat), num_target="1"),
        objective=dict(name="multi:softprob")), version=[3, 0, 4])
