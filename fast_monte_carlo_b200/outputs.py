"""Outputs the consumer (`edge_finder.py`) reads, computed either from the per-game score table or
directly from the joint (points A, points B) histogram the kernels accumulate.

`edge_finder.game_market_odds` / `moneyline_from_sims` (edge_finder.py:235-336) only use the joint
distribution of (pts, opp_pts) per orientation, so every statistic they print is a function of the
histogram; `scores_*.csv` can still be materialised for a literal drop-in.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import pandas as pd

HIST_BINS = 128


def _chunks(n: int, workers: int):
    """Even-aligned row ranges for the worker threads (numpy copies release the GIL)."""
    step = max(2, ((n + workers - 1) // workers + 1) & ~1)
    return [(lo, min(n, lo + step)) for lo in range(0, n, step)]


def _pool(n: int):
    import os
    from concurrent.futures import ThreadPoolExecutor
    try:
        cpus = len(os.sched_getaffinity(0))
    except AttributeError:          # pragma: no cover
        cpus = os.cpu_count() or 1
    w = max(1, min(16, cpus, n // 500_000))
    return (ThreadPoolExecutor(max_workers=w) if w > 1 else None), w


def _alternating_names(first: str, second: str, n: int, pool=None, workers: int = 1):
    """Arrow large_string array `first, second, first, ...` of length n, assembled from its buffers (no Python string
    objects: a 10 M-row column takes tens of milliseconds instead of seconds)."""
    import pyarrow as pa
    fb, sb = first.encode("utf-8"), second.encode("utf-8")
    la, lb = len(fb), len(sb)
    offsets = np.empty(n + 1, dtype=np.int64)
    total = (n // 2) * (la + lb) + (n & 1) * la
    data = np.empty(total, dtype=np.uint8)
    pair = np.frombuffer(fb + sb, dtype=np.uint8)

    def fill(lo, hi):          # lo is even
        m = hi - lo
        base = (lo // 2) * (la + lb)
        k = np.arange(0, (m + 1) // 2 + 1, dtype=np.int64) * (la + lb) + base
        offsets[lo:hi + 1:2] = k[: (m + 2) // 2]
        offsets[lo + 1:hi + 1:2] = k[: (m + 1) // 2] + la
        nb = int(offsets[hi]) - base
        full = nb // (la + lb)
        if full:
            data[base:base + full * (la + lb)].reshape(full, la + lb)[:] = pair
        if nb - full * (la + lb):
            data[base + full * (la + lb):base + nb] = pair[: nb - full * (la + lb)]

    parts = _chunks(n, workers) if n else []
    if pool is not None and len(parts) > 1:
        list(pool.map(lambda r: fill(*r), parts))
    else:
        for r in parts:
            fill(*r)
    if n == 0:
        offsets[0] = 0
    return pa.LargeStringArray.from_buffers(n, pa.py_buffer(offsets), pa.py_buffer(data))


def name_columns(team_a: str, team_b: str, n: int, first_game: int = 0):
    """The `team` / `opp` columns of `sims_frame` for games [first_game, first_game + n): they depend on the names and
    the game parity only, so `api.simulate_matchup` builds them on a worker thread WHILE the kernel runs."""
    n = int(n)
    p = int(first_game) & 1
    names = (team_a, team_b) if p == 0 else (team_b, team_a)
    dt = pd.Series(["x"]).dtype        # what `pd.DataFrame(rows)` of the reference gives its name columns
    if dt == object:
        nm = np.array(names, dtype=object)
        idx = np.arange(n) & 1
        return nm[idx], nm[1 - idx]
    pool, workers = _pool(n)
    try:
        team = pd.array(_alternating_names(names[0], names[1], n, pool, workers), dtype=dt)
        other = pd.array(_alternating_names(names[1], names[0], n, pool, workers), dtype=dt)
    finally:
        if pool is not None:
            pool.shutdown()
    return team, other


def sims_frame(team_a: str, team_b: str, scores: np.ndarray, first_game: int = 0, names=None) -> pd.DataFrame:
    """Per-game table with the reference's columns and row order (FMC:1501-1509): game g has the
    opening-kickoff receiver as `team`; even games are A-first, odd games B-first.  The two name columns have
    the dtype pandas infers for the reference's own lists of names; they are assembled from Arrow buffers and the
    point columns by strided copies, on a few threads (10 M rows: ~0.1 s instead of 2.5 s through 20 M Python
    string objects).  `names` = a `name_columns(...)` result computed ahead of time."""
    n = int(scores.shape[0])
    p = int(first_game) & 1
    pts = np.empty(n, dtype=np.int64)
    opp = np.empty(n, dtype=np.int64)
    pool, workers = _pool(n)

    def points(lo, hi):        # rows lo, lo + 2, ... (lo even) have parity p: A first when p == 0
        pts[lo:hi:2] = scores[lo:hi:2, p]; opp[lo:hi:2] = scores[lo:hi:2, 1 - p]
        pts[lo + 1:hi:2] = scores[lo + 1:hi:2, 1 - p]; opp[lo + 1:hi:2] = scores[lo + 1:hi:2, p]

    try:
        parts = _chunks(n, workers) if n else []
        if pool is not None and len(parts) > 1:
            list(pool.map(lambda r: points(*r), parts))
        else:
            for r in parts:
                points(*r)
    finally:
        if pool is not None:
            pool.shutdown()
    team, other = names if names is not None else name_columns(team_a, team_b, n, first_game)
    return pd.DataFrame({"team": team, "opp": other, "pts": pts, "opp_pts": opp}, copy=False)


def summary_frame(sims_df: pd.DataFrame) -> pd.DataFrame:
    """groupby(team) mean/sd/win_rate exactly as FMC:1681-1687 (sample std, ties are not wins)."""
    win = (sims_df["pts"].to_numpy() > sims_df["opp_pts"].to_numpy())
    tmp = sims_df.assign(_win=win)
    out = tmp.groupby("team").agg(
        mean_pts=("pts", "mean"), sd_pts=("pts", "std"),
        mean_opp=("opp_pts", "mean"), sd_opp=("opp_pts", "std"),
        win_rate=("_win", "mean"))
    return out


def histogram_from_scores(scores: np.ndarray, first_game: int = 0) -> np.ndarray:
    """[2, BINS, BINS] joint histogram (orientation = game parity) from a per-game (A, B) table."""
    h = np.zeros((2, HIST_BINS, HIST_BINS), dtype=np.int64)
    g = np.arange(first_game, first_game + scores.shape[0]) & 1
    a = np.minimum(scores[:, 0], HIST_BINS - 1)
    b = np.minimum(scores[:, 1], HIST_BINS - 1)
    np.add.at(h, (g, a, b), 1)
    return h


def _oriented(hist2: np.ndarray, team_is_a: bool):
    """(pts, opp_pts, weight) grids of the orientation in which `team` received the kickoff --
    the only rows edge_finder.game_market_odds keeps (edge_finder.py:301-302)."""
    o = 0 if team_is_a else 1
    w = hist2[o].astype(np.float64)
    a = np.arange(HIST_BINS)[:, None] + np.zeros((1, HIST_BINS), dtype=np.int64)
    b = np.arange(HIST_BINS)[None, :] + np.zeros((HIST_BINS, 1), dtype=np.int64)
    return (a, b, w) if team_is_a else (b, a, w)


def _weighted_median(values: np.ndarray, weights: np.ndarray) -> float:
    """np.median of the multiset (even counts average the two middle values)."""
    order = np.argsort(values, kind="stable")
    v = values[order]
    w = weights[order]
    keep = w > 0
    v, w = v[keep], w[keep]
    c = np.cumsum(w)
    n = c[-1]
    lo = v[np.searchsorted(c, (n + 1) // 2, side="left")]
    hi = v[np.searchsorted(c, n // 2 + 1, side="left")]
    return float(lo + hi) / 2.0 if n % 2 == 0 else float(lo)


def prob_to_american(p: float) -> int:
    """edge_finder._prob_to_american (edge_finder.py:70-75)."""
    p = float(np.clip(p, 1e-6, 1 - 1e-6))
    return int(round(-100 * p / (1 - p))) if p >= 0.5 else int(round(100 * (1 - p) / p))


def moneyline_from_hist(hist2: np.ndarray, team: str, opp: str) -> Dict[str, Dict]:
    """edge_finder.moneyline_from_sims (edge_finder.py:249-281) with `team` = A: p_team from the
    A-first orientation, p_opp from the B-first one (they need not sum to 1)."""
    a, b, w = _oriented(hist2, True)
    p_team = float((w * (a > b)).sum() / w.sum())
    b2, a2, w2 = _oriented(hist2, False)
    p_opp = float((w2 * (b2 > a2)).sum() / w2.sum())
    return {"team": {"name": team, "p_win": round(p_team, 6), "ml_fair": prob_to_american(p_team)},
            "opp": {"name": opp, "p_win": round(p_opp, 6), "ml_fair": prob_to_american(p_opp)}}


def game_market_odds_from_hist(hist2: np.ndarray, team: str, opp: str, *, team_is_a: bool = True,
                               spread: Optional[float] = None, total: Optional[float] = None) -> Dict[str, Dict]:
    """edge_finder.game_market_odds (edge_finder.py:283-336) from the histogram."""
    pts, opp_pts, w = _oriented(hist2, team_is_a)
    n = w.sum()
    if n <= 0:
        raise ValueError("No rows from the TEAM perspective found in scores file.")
    out: Dict[str, Dict] = {}
    if spread is not None:
        margin = (pts - opp_pts).astype(np.float64)
        tgt = -float(spread)
        p_cover = float((w * (margin > tgt)).sum() / n)
        p_not = float((w * (margin < tgt)).sum() / n)
        p_push = float((w * np.isclose(margin, tgt, atol=1e-9)).sum() / n)
        out["spread"] = {
            "team": team, "opp": opp, "spread": float(spread), "samples": int(n),
            "p_cover": round(p_cover, 6), "p_notcover": round(p_not, 6), "push_rate": round(p_push, 6),
            "american_cover": prob_to_american(p_cover), "american_notcover": prob_to_american(p_not),
            "mean_margin": float((w * margin).sum() / n),
            "median_margin": _weighted_median(margin.ravel(), w.ravel()),
        }
    if total is not None:
        totals = (pts + opp_pts).astype(np.float64)
        T = float(total)
        p_over = float((w * (totals > T)).sum() / n)
        p_under = float((w * (totals < T)).sum() / n)
        p_push = float((w * np.isclose(totals, T, atol=1e-9)).sum() / n)
        out["total"] = {
            "team": team, "opp": opp, "total": float(total), "samples": int(n),
            "p_over": round(p_over, 6), "p_under": round(p_under, 6), "push_rate": round(p_push, 6),
            "american_over": prob_to_american(p_over), "american_under": prob_to_american(p_under),
            "mean_total": float((w * totals).sum() / n),
            "median_total": _weighted_median(totals.ravel(), w.ravel()),
        }
    if not out:
        raise ValueError("Provide at least one of spread= or total=.")
    return out


def summary_from_hist(hist2: np.ndarray, team_a: str, team_b: str) -> pd.DataFrame:
    """The `summary` frame of simulate_upcoming_matchup (FMC:1681-1687) from the histogram: each
    team's row covers the orientation in which it received the opening kickoff."""
    rows = {}
    for name, is_a in ((team_a, True), (team_b, False)):
        pts, opp_pts, w = _oriented(hist2, is_a)
        n = w.sum()
        def mean_sd(x):
            m = (w * x).sum() / n
            var = (w * (x - m) ** 2).sum() / (n - 1) if n > 1 else np.nan
            return float(m), float(np.sqrt(var))
        mp, sp_ = mean_sd(pts.astype(np.float64))
        mo, so = mean_sd(opp_pts.astype(np.float64))
        rows[name] = dict(mean_pts=mp, sd_pts=sp_, mean_opp=mo, sd_opp=so,
                          win_rate=float((w * (pts > opp_pts)).sum() / n))
    df = pd.DataFrame.from_dict(rows, orient="index")
    df.index.name = "team"
    return df.sort_index()


# ---------------------------------------------------------------------------------------------
# literal score tables at scale: the file edge_finder.py opens
# ---------------------------------------------------------------------------------------------
def write_scores_table(path: str, team_a: str, team_b: str, scores: np.ndarray, first_game: int = 0,
                       chunk_rows: int = 4_000_000) -> int:
    """Writes `scores_<base>.parquet|csv` (columns team, opp, pts, opp_pts; rows alternate A-first /
    B-first exactly like FMC:1501-1509) straight from the per-game (A, B) score array, in row-group
    chunks, with the two team-name columns dictionary-encoded -- a 10 M-game table is ~25 MB of parquet
    and never exists as a pandas object frame.  Returns the number of rows written."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    n = int(scores.shape[0])
    names = pa.array([team_a, team_b], type=pa.string())
    schema = pa.schema([("team", pa.dictionary(pa.int8(), pa.string())), ("opp", pa.dictionary(pa.int8(), pa.string())),
                        ("pts", pa.int64()), ("opp_pts", pa.int64())])
    is_parquet = path.lower().endswith(".parquet")
    writer = pq.ParquetWriter(path, schema, compression="zstd") if is_parquet else None
    try:
        for lo in range(0, max(n, 1), chunk_rows):
            sc = scores[lo:lo + chunk_rows]
            g = np.arange(first_game + lo, first_game + lo + sc.shape[0])
            b_first = (g & 1).astype(np.int8)
            pts = np.where(b_first == 0, sc[:, 0], sc[:, 1]).astype(np.int64)
            opp = np.where(b_first == 0, sc[:, 1], sc[:, 0]).astype(np.int64)
            tbl = pa.table({
                "team": pa.DictionaryArray.from_arrays(pa.array(b_first), names),
                "opp": pa.DictionaryArray.from_arrays(pa.array((1 - b_first).astype(np.int8)), names),
                "pts": pa.array(pts), "opp_pts": pa.array(opp)}, schema=schema)
            if is_parquet:
                writer.write_table(tbl)
            else:
                import pyarrow.csv as pacsv
                cols = {c: (tbl[c].cast(pa.string()) if c in ("team", "opp") else tbl[c]) for c in tbl.column_names}
                plain = pa.table(cols)
                with open(path, "ab" if lo else "wb") as fh:
                    pacsv.write_csv(plain, fh, write_options=pacsv.WriteOptions(include_header=(lo == 0)))
    finally:
        if writer is not None:
            writer.close()
    return n
