#!/usr/bin/env python
"""How many DISTINCT (family, orientation, threshold-rank vector) keys does a run of configs[1] visit?

A tree ensemble's output is a function of which side of every split threshold each feature lies on.  On the
per-orientation specialised forests (constants folded exactly as csrc/fmc_pack.hpp folds them) a request's
varying features reduce to a small vector of per-feature threshold ranks, so two requests with the same key take
the same branch at every node and get bit-identical outputs: a memo keyed on it is EXACT (unlike the reference's
own caches, FMC:68-94, which bin coarsely).  This script measures, with the CPU oracle's per-iteration trace
(test infrastructure, never the product path), the number of distinct keys per family as games accumulate --
the hit rate an infinite memo would have.  Keys are computed for EVERY play state for every family, which is a
superset of the requests actually made, so the hit rates printed are lower bounds.

    python scripts/memo_study.py [--games 200000] [--chunk 20000]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fast_monte_carlo_b200 import artifacts as art, synth  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

KSU = (15.6, 35.7, 20.0)
ISU = (11.0, 31.5, 20.6)
FAMS = ("pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards")
VARYING = (0, 1, 2, 3, 4, 5, 12, 13, 14, 15, 16)


def reachable_thresholds(f, fold, active):
    """Thresholds per varying numeric on the forest specialised on `fold` (numeric k -> constant) and the hot
    one-hot columns `active` -- the same folding as fmc_pack.hpp Builder::build."""
    thr = {k: set() for k in VARYING}
    zm = f.kind == art.KIND_XGB and f.zero_is_missing

    def const_left(i, v):
        if f.kind == art.KIND_XGB:
            if zm and v == 0.0:
                return bool(f.default_left[i])
            return np.float32(v) < f.thr[i]
        return np.float32(v) <= f.thr[i]

    for root in f.tree_root:
        stack = [int(root)]
        while stack:
            i = stack.pop()
            while f.left[i] >= 0:
                col = int(f.feat[i])
                if f.num_base <= col < f.num_base + f.n_num:
                    k = col - f.num_base
                    if k in fold:
                        i = int(f.left[i] if const_left(i, fold[k]) else f.right[i])
                        continue
                    thr[k].add(float(f.thr[i]))
                    stack.append(int(f.right[i]))
                    i = int(f.left[i])
                else:
                    v = 1.0 if col in active else 0.0
                    i = int(f.left[i] if const_left(i, v) else f.right[i])
    return {k: np.array(sorted(v), dtype=np.float32) for k, v in thr.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=200_000)
    ap.add_argument("--chunk", type=int, default=20_000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    ms = synth.with_synthetic_stage2(art.load_default_models())
    co.load_models(ms)
    cfg = co.make_config(ms, KSU, ISU, stage2="booster")
    sp = (KSU, ISU)
    tables = {}
    for name in FAMS:
        f = ms[name]
        active = {g.column_of("Unknown") for g in f.groups[:2]} - {-1}
        for off in (0, 1):
            de = off ^ 1
            fold = {6: 3.0, 7: 3.0, 8: sp[off][0], 9: sp[off][1], 10: sp[de][2], 11: sp[de][0]}
            tables[(name, off)] = reachable_thresholds(f, fold, active)
            t = tables[(name, off)]
            bits = sum(int(np.ceil(np.log2(len(v) + 2))) for v in t.values() if len(v))
            print(name, off, {art.NUM_FEATURES[k]: len(v) for k, v in t.items() if len(v)}, "bits", bits, flush=True)

    seen = {k: set() for k in tables}
    rows = []
    tot = {k: 0 for k in ("plays", "pass", "comp", "run", "sack")}
    done = 0
    while done < args.games:
        n = min(args.chunk, args.games - done)
        r = co.simulate(cfg, n, game0=done, seed=20251018, trace=True)
        for k in tot:
            tot[k] += r["counters"][k]
        tr = r["trace"].reshape(-1, 8)
        tr = tr[~np.isnan(tr[:, 1])]
        game_first = None
        off_is_first = tr[:, 0] == 1.0
        # orientation = team index on offense: game g's first receiver is team (g & 1); the trace does not carry the
        # game id per row, so rebuild it
        iters = r["iters"]
        gid = np.repeat(np.arange(done, done + n), np.minimum(iters, co.MAX_ITERS))
        first = (gid & 1).astype(np.int64)
        team = np.where(off_is_first, first, first ^ 1)
        down = tr[:, 1]
        sec = tr[:, 2]
        sd = np.where(off_is_first, tr[:, 3] - tr[:, 4], tr[:, 4] - tr[:, 3])
        dist = tr[:, 5]
        ytg = tr[:, 6]
        feats = {0: down, 1: dist, 2: ytg, 3: (ytg <= 20.0) * 1.0, 4: sd, 5: sec,
                 12: (dist >= ytg - 0.5) * 1.0, 13: ((down == 4) & (dist <= 2.0)) * 1.0, 14: (ytg <= 33.0) * 1.0,
                 15: np.where(sec > 1800, 1.0, 2.0), 16: ((sec % 1800) <= 120) * 1.0}
        feats = {k: v.astype(np.float32) for k, v in feats.items()}
        for (name, off), t in tables.items():
            m = team == off
            f = ms[name]
            xgb = f.kind == art.KIND_XGB
            zm = xgb and f.zero_is_missing
            key = np.zeros(int(m.sum()), dtype=np.uint64)
            for k in VARYING:
                if len(t[k]) == 0:
                    continue
                x = feats[k][m]
                rk = np.searchsorted(t[k], x, side="right" if xgb else "left").astype(np.uint64)
                radix = len(t[k]) + 1
                if zm:
                    rk = np.where(x == 0.0, np.uint64(radix), rk)
                    radix += 1
                key = key * np.uint64(radix) + rk
            u = np.unique(key)
            seen[(name, off)].update(u.tolist())
        done += n
        row = {"games": done, **{f"{nm}:{o}": len(s) for (nm, o), s in seen.items()}, **tot}
        rows.append(row)
        req = {"pass_stage1": tot["pass"], "pass_stage2": tot["pass"] - tot["comp"], "pass_yards": tot["comp"],
               "run_yards": tot["run"], "sack_yards": tot["sack"]}
        msg = [f"games {done}"]
        for name in FAMS:
            d = len(seen[(name, 0)]) + len(seen[(name, 1)])
            msg.append(f"{name} {d} keys / {req[name]} req = hit>= {1 - d / max(req[name], 1):.4f}")
        print(" | ".join(msg), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
