"""Usage tables: who throws, runs and is targeted, and which of them get a box-score line.

Host-side mirror of the reference's player plumbing:

  focus sheet      `2025_week1_players.csv` (team, player, pos, usage, stat, yards) -> per team three
                   share tables + the "track" sets            `_build_focus_usage_tables` FMC:511-602
  fallback files   `usage_{qb,rush,target}_share.csv`          `_load_usage_table` FMC:487-505
  resolution       focus first, else files, else "Unknown"     `_usage_from_focus_or_fallback` FMC:228-249

The reference samples a name per play with `Generator.choice(len(df), p=df['share'])` (FMC:625-635),
feeds it to the models as `passer_name` / `target_name` / `rusher_name` (FMC:1079-1081, 1216; the
synthetic remainder receiver `__Other__` is fed as "Unknown", FMC:1066) and keeps a per-game box line
only for names in the team's track sets (FMC:1062-1063, 1204).  `resolve_team` reduces one team's
tables to exactly that: shares, output slots of the tracked names, and the one-hot column each name
lights in every model (the OneHotEncoder(handle_unknown='ignore') half of the preprocessors).
It runs once per team; sampling, one-hots and the box scores themselves run in the CUDA kernel.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import artifacts as art

OTHER_SENTINEL = "__Other__"            # FMC:509
FOCUS_PLAYERS_CSV = "2025_week1_players.csv"   # FMC:508
ROLES = ("pass", "rush", "rec")
ROLE_NAME_COL = {"pass": "passer_name", "rush": "rusher_name", "rec": "receiver_name"}
ROLE_LABEL = {"pass": "QB", "rush": "Rusher", "rec": "Receiver"}      # FMC:1275, 1284, 1293
_FOCUS_STAT = {"pass": "pass_yards", "rush": "rush_yards", "rec": "rec_yards"}
_FILES = {"pass": "usage_qb_share.csv", "rush": "usage_rush_share.csv", "rec": "usage_target_share.csv"}
MAX_USAGE = 32       # entries per table the kernels hold (include/fmc.h FMC_MAX_USAGE)
MAX_NAME_ROWS = {"pass": 4, "rush": 8, "rec": 8}   # names per role that some model has a one-hot column for
# which name group of which model a role feeds (model name -> group name); FMC:1079-1081, 1216
_ROLE_GROUP = {"pass": "passer_name", "rush": "rusher_name", "rec": "target_name"}
_PLAYER_MODELS = ("pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards")


def _unknown(col: str) -> pd.DataFrame:
    return pd.DataFrame({col: ["Unknown"], "share": [1.0]})


def _share_table(rows: Optional[pd.DataFrame], col: str) -> pd.DataFrame:
    """One stat bucket of the focus sheet -> (name, share) summing to 1 (`_mk`, FMC:534-581).

    Percentages become fractions, duplicate players are summed, a total below 1 leaves the remainder
    to the synthetic `__Other__` entry, anything unusable degrades to the single "Unknown" row.
    """
    if rows is None or rows.empty:
        return _unknown(col)
    t = rows[["player", "usage"]].copy()
    t["usage"] = pd.to_numeric(t["usage"], errors="coerce").fillna(0.0).clip(lower=0.0)
    if t["usage"].max() > 1.5:
        t["usage"] = t["usage"] / 100.0
    t = t.groupby("player", as_index=False)["usage"].sum()
    total = float(t["usage"].sum())
    if not np.isfinite(total) or total <= 0.0:
        return _unknown(col)
    if total >= 1.0 - 1e-9:
        t["share"] = t["usage"] / total
    else:
        t["share"] = t["usage"]
        rest = 1.0 - float(t["share"].sum())
        if rest > 1e-12:
            t.loc[len(t)] = {"player": OTHER_SENTINEL, "usage": rest, "share": rest}
    t = t.rename(columns={"player": col})[[col, "share"]]
    total = float(t["share"].sum())
    if not np.isfinite(total) or total <= 0.0:
        return _unknown(col)
    t["share"] = (t["share"] / total).clip(lower=0.0)
    return t


def build_focus_usage_tables(path: str = FOCUS_PLAYERS_CSV) -> Dict[str, dict]:
    """team -> dict(qb_df, ru_df, tg_df, track_pass, track_rush, track_rec); {} when the sheet is absent."""
    if not os.path.exists(path):
        return {}
    df = pd.read_csv(path)
    df["team"] = df["team"].astype(str).str.strip()
    df["player"] = df["player"].astype(str).str.strip()
    df["pos"] = df["pos"].astype(str).str.upper().str.strip()
    df["stat"] = df["stat"].astype(str).str.strip().str.lower()
    df["usage"] = pd.to_numeric(df["usage"], errors="coerce")
    out: Dict[str, dict] = {}
    for team, g in df.groupby("team"):
        bucket = {r: g[g["stat"] == _FOCUS_STAT[r]][["player", "usage"]].copy() for r in ROLES}
        out[team] = dict(
            qb_df=_share_table(bucket["pass"], "passer_name"),
            ru_df=_share_table(bucket["rush"], "rusher_name"),
            tg_df=_share_table(bucket["rec"], "receiver_name"),
            track_pass=set(bucket["pass"]["player"].astype(str)),
            track_rush=set(bucket["rush"]["player"].astype(str)),
            track_rec=set(bucket["rec"]["player"].astype(str)))
    return out


def load_usage_table(path: str, team: str, year: int, who_col: str) -> Optional[pd.DataFrame]:
    """One `usage_*_share.csv` filtered to (team, year), shares clipped at 0 and renormalised (FMC:487-505)."""
    try:
        df = pd.read_csv(path)
        df = df[(df["offense"] == team) & (df["year"] == year)].copy()
        if df.empty or who_col not in df.columns:
            return None
        df = df[[who_col, "share"]].dropna()
        if df.empty:
            return None
        s = df["share"].clip(lower=0)
        s = s / s.sum() if s.sum() > 0 else pd.Series([1.0], index=[0])
        df["share"] = s.values
        return df
    except Exception:
        return None


def usage_for_team(team: str, year: int, focus: Optional[Dict[str, dict]] = None, directory: str = "."):
    """(qb_df, ru_df, tg_df, track_pass, track_rush, track_rec) -- focus sheet first (FMC:228-249)."""
    if focus and team in focus:
        f = focus[team]
        return (f["qb_df"].copy(), f["ru_df"].copy(), f["tg_df"].copy(),
                f["track_pass"], f["track_rush"], f["track_rec"])
    tabs = []
    for r in ROLES:
        t = load_usage_table(os.path.join(directory, _FILES[r]), team, year, ROLE_NAME_COL[r])
        tabs.append(t if t is not None else _unknown(ROLE_NAME_COL[r]))
    return tabs[0], tabs[1], tabs[2], set(), set(), set()


def py_round1(x: np.ndarray) -> np.ndarray:
    """Python's `round(v, 1)` (correctly rounded on the decimal value, FMC:1276, 1286, 1296) for an array:
    NumPy's round scales by 10 first and differs on values such as 1.15."""
    x = np.asarray(x, dtype=np.float64)
    out = np.round(x, 1)
    t = np.abs(x) * 10.0
    near = np.abs((t - np.floor(t)) - 0.5) < 1e-6
    for i in np.nonzero(near)[0]:
        out[i] = round(float(x[i]), 1)
    return out


# ---------------------------------------------------------------------------------------------
@dataclass
class RoleUsage:
    names: List[str]
    share: np.ndarray                 # f64, as handed to Generator.choice
    slot: List[int]                   # output slot of a tracked name, -1 otherwise
    col: Dict[str, List[int]]         # model name -> hot column per entry (-1: not a category)


def name_rows(ru: "RoleUsage") -> List[int]:
    """0/1 feature row of every usage entry, -1 for names no model has a column for (the same numbering
    fmc_set_usage derives, csrc/fmc_abi.cu)."""
    out, nxt = [], 0
    for e in range(len(ru.names)):
        if any(ru.col[m][e] >= 0 for m in ru.col):
            out.append(nxt)
            nxt += 1
        else:
            out.append(-1)
    return out


@dataclass
class TeamUsage:
    role: Dict[str, RoleUsage]
    slots: List[Tuple[str, str]] = field(default_factory=list)    # slot -> (role, name)

    @property
    def trivial(self) -> bool:
        """The shipped configuration: one "Unknown" per role, nothing tracked (FMC:246-249)."""
        return not self.slots and all(r.names == ["Unknown"] for r in self.role.values())


def _model_name(role: str, name: str) -> str:
    # FMC:1066: only the RECEIVER remainder is renamed for the models; `__Other__` passers / rushers go in as is
    return "Unknown" if (role == "rec" and name == OTHER_SENTINEL) else name


def resolve_team(tc, models: art.ModelSet) -> TeamUsage:
    """TeamContext (qb_share / rush_share / target_share + track sets) -> what the kernels read."""
    tables = {"pass": tc.qb_share, "rush": tc.rush_share, "rec": tc.target_share}
    tracks = {"pass": tc.track_pass, "rush": tc.track_rush, "rec": tc.track_rec}
    out = TeamUsage(role={})
    for r in ROLES:
        df = tables[r] if tables[r] is not None else _unknown(ROLE_NAME_COL[r])
        names = [str(x) for x in df[ROLE_NAME_COL[r]].tolist()]
        share = np.asarray(df["share"].values, dtype=np.float64)
        if not (1 <= len(names) <= MAX_USAGE):
            raise ValueError(f"{tc.name}: {len(names)} {r} usage entries; the kernels hold 1..{MAX_USAGE}")
        if not np.all(np.isfinite(share)) or np.any(share < 0) or not share.sum() > 0:
            raise ValueError(f"{tc.name}: {r} shares must be finite, non-negative and not all zero")
        track = tracks[r] or set()
        slot = []
        for nm in names:
            if nm in track and nm != OTHER_SENTINEL:     # `__Other__` never gets a line (FMC:1272, 1281, 1290)
                slot.append(len(out.slots))
                out.slots.append((r, nm))
            else:
                slot.append(-1)
        col: Dict[str, List[int]] = {}
        for m in _PLAYER_MODELS:
            g = models[m].group(_ROLE_GROUP[r]) if m in models else None
            col[m] = [g.column_of(_model_name(r, nm)) if g is not None else -1 for nm in names]
            hot = [c for c in col[m] if c >= 0]
            if len(hot) != len(set(hot)):
                raise ValueError(f"{tc.name}: two {r} usage entries map to the same {m} column")
        known = sum(1 for e in range(len(names)) if any(col[m][e] >= 0 for m in col))
        if known > MAX_NAME_ROWS[r]:
            raise ValueError(f"{tc.name}: {known} {r} names that the models have one-hot columns for; "
                             f"the kernels hold {MAX_NAME_ROWS[r]} (names unknown to every model are not limited)")
        out.role[r] = RoleUsage(names=names, share=share, slot=slot, col=col)
    return out


def player_rows(box, game0: int, team_names: Sequence[str], usage: Sequence[TeamUsage]) -> pd.DataFrame:
    """Per-game box `[games][2][n_slots][6]` (yds, att|tgt, comp|rec, td, INT, sacks) -> the reference's
    players table (`flatten_player_box_rows`, PLAYER_COLS, FMC:1259-1299): one row per game, team and
    tracked name that was sampled at least once in that game; `sim` = game id, `start` = "A"/"B".
    Row order: game, receiving team first, then QB / Rusher / Receiver in usage-table order (the reference
    orders names by first appearance inside a game)."""
    from .api import PLAYER_COLS
    from .native import box_slot
    n = box.shape[0]
    frames = []
    gid = np.arange(game0, game0 + n, dtype=np.int64)
    first = (gid & 1).astype(np.int64)                     # team that received the opening kickoff
    for t in (0, 1):
        for s, (role, name) in enumerate(usage[t].slots):
            rec = box_slot(box, t, s)
            seen = (rec[:, 1] > 0) | (rec[:, 5] > 0)        # _ensure_player ran: a pass call, a target, a carry
            if not seen.any():
                continue
            k = np.nonzero(seen)[0]
            z = np.zeros(k.shape[0], dtype=np.int64)
            zf = np.zeros(k.shape[0], dtype=np.float64)
            yds = py_round1(rec[k, 0])
            cnt = lambda j: rec[k, j].astype(np.int64)
            d = dict(sim=gid[k], start=np.where(first[k] == 0, "A", "B"), team=team_names[t], opp=team_names[t ^ 1],
                     player=name, role=ROLE_LABEL[role],
                     pass_att=z, pass_comp=z, pass_yds=zf, pass_td=z, INT=z, sacks=z,
                     rush_att=z, rush_yds=zf, rush_td=z, rec=z, tgt=z, rec_yds=zf, rec_td=z)
            if role == "pass":
                d.update(pass_att=cnt(1), pass_comp=cnt(2), pass_yds=yds, pass_td=cnt(3), INT=cnt(4), sacks=cnt(5))
            elif role == "rush":
                d.update(rush_att=cnt(1), rush_yds=yds, rush_td=cnt(3))
            else:
                d.update(tgt=cnt(1), rec=cnt(2), rec_yds=yds, rec_td=cnt(3))
            f = pd.DataFrame(d)
            f["_team_order"] = np.where(first[k] == t, 0, 1)
            f["_slot"] = s
            frames.append(f)
    if not frames:
        return pd.DataFrame(columns=PLAYER_COLS)
    out = pd.concat(frames, ignore_index=True)
    out = out.sort_values(["sim", "_team_order", "_slot"], kind="stable").drop(columns=["_team_order", "_slot"])
    return out[PLAYER_COLS].reset_index(drop=True)


_STAT_FIELD = {  # players_* column -> (role, box field); PLAYER_COLS FMC:1259-1264
    "pass_att": ("pass", 1), "pass_comp": ("pass", 2), "pass_yds": ("pass", 0), "pass_td": ("pass", 3),
    "INT": ("pass", 4), "sacks": ("pass", 5),
    "rush_att": ("rush", 1), "rush_yds": ("rush", 0), "rush_td": ("rush", 3),
    "tgt": ("rec", 1), "rec": ("rec", 2), "rec_yds": ("rec", 0), "rec_td": ("rec", 3),
}
_STAT_ALIASES = {"pass_yards": "pass_yds", "rush_yards": "rush_yds", "rec_yards": "rec_yds"}   # edge_finder.py:12-17


def player_prop_odds_from_box(box, team_names: Sequence[str], usage: Sequence[TeamUsage],
                              team: str, player: str, stat: str, line: float) -> Dict[str, object]:
    """`edge_finder.player_prop_odds` (edge_finder.py:168-231) straight from the per-game box, without
    materialising `players_*`: over/under/push rates of one player's stat against `line`, fair American
    odds, mean / median / p75 / p90 and the better side at -110.  Like the players table, only games in
    which the name was sampled at least once count as samples."""
    from .outputs import prob_to_american
    col = _STAT_ALIASES.get(stat, stat)
    if col not in _STAT_FIELD:
        raise ValueError(f"Stat '{stat}' (mapped to '{col}') not present in the player box.")
    role, fld = _STAT_FIELD[col]
    t = [i for i, nm in enumerate(team_names) if nm.lower() == team.lower()]
    hit = [s for s, (r, nm) in enumerate(usage[t[0]].slots) if r == role and nm.lower() == player.lower()] if t else []
    if not hit:
        raise ValueError(f"No rows found for {player} on {team}.")
    from .native import box_slot
    rec = box_slot(box, t[0], hit[0])
    seen = (rec[:, 1] > 0) | (rec[:, 5] > 0)
    vals = rec[seen, fld]
    if fld == 0:
        vals = py_round1(vals)
    if vals.size == 0:
        raise ValueError(f"No rows found for {player} on {team}.")
    p_over = float(np.mean(vals > line))
    p_under = float(np.mean(vals < line))
    p_push = float(np.mean(np.isclose(vals, line, atol=1e-9)))
    ev = lambda p: p * (100.0 * (100.0 / 110.0)) - (1.0 - p) * 100.0      # EV per $100 at -110
    implied = 110.0 / 210.0
    side, best_ev, edge = (("Over", ev(p_over), p_over - implied) if ev(p_over) >= ev(1.0 - p_over)
                           else ("Under", ev(1.0 - p_over), (1.0 - p_over) - implied))
    return {
        "team": team, "player": player, "role": ROLE_LABEL[role], "stat": col, "line": float(line),
        "samples": int(vals.size), "p_over": round(p_over, 4), "p_under": round(p_under, 4),
        "push_rate": round(p_push, 4), "american_over": prob_to_american(p_over),
        "american_under": prob_to_american(p_under), "mean": float(np.mean(vals)), "median": float(np.median(vals)),
        "p75": float(np.percentile(vals, 75)), "p90": float(np.percentile(vals, 90)),
        "best_side": side, "edge": round(edge * 100, 2), "ev_per_$100": round(best_ev, 2),
    }


# ---------------------------------------------------------------------------------------------
# per-player histograms (include/fmc.h FMC_PH_*): what the kernel accumulates with atomics and ranks merge with
# one integer all-reduce -- the scalable form of `players_*`
# ---------------------------------------------------------------------------------------------
PH_YDS_BINS, PH_YDS_OFFSET, PH_CNT_BINS = 8192, 1000, 128
PH_BINS = PH_YDS_BINS + 5 * PH_CNT_BINS


def player_hist_from_box(box, usage: Sequence[TeamUsage]) -> np.ndarray:
    """Host restatement of the kernel's histogram step: uint32[2][n_slots][PH_BINS] from a per-game box; a game
    counts for a line only if the name was sampled in it.  Yards are rounded like Python's round(x, 1)."""
    from .native import box_slot
    n_slots = box.shape[2]
    h = np.zeros((2, n_slots, PH_BINS), dtype=np.uint32)
    for t in (0, 1):
        for s in range(len(usage[t].slots)):
            rec = box_slot(box, t, s)
            seen = (rec[:, 1] > 0) | (rec[:, 5] > 0)
            if not seen.any():
                continue
            tenths = np.rint(py_round1(rec[seen, 0]) * 10.0).astype(np.int64) + PH_YDS_OFFSET
            h[t, s, :PH_YDS_BINS] = np.bincount(np.clip(tenths, 0, PH_YDS_BINS - 1), minlength=PH_YDS_BINS)
            for k in range(5):
                c = np.minimum(rec[seen, 1 + k].astype(np.int64), PH_CNT_BINS - 1)
                lo = PH_YDS_BINS + k * PH_CNT_BINS
                h[t, s, lo:lo + PH_CNT_BINS] = np.bincount(c, minlength=PH_CNT_BINS)
    return h


def _percentile_from_counts(values: np.ndarray, counts: np.ndarray, q: float) -> float:
    """np.percentile(sample, q) (linear interpolation) of the sample a histogram stands for."""
    cum = np.cumsum(counts)
    n = int(cum[-1])
    pos = (n - 1) * (q / 100.0)
    lo, hi = int(np.floor(pos)), int(np.ceil(pos))
    a = float(values[np.searchsorted(cum, lo + 1, side="left")])
    b = float(values[np.searchsorted(cum, hi + 1, side="left")])
    t = pos - lo
    return b - (b - a) * (1.0 - t) if t >= 0.5 else a + (b - a) * t      # numpy's _lerp


def player_prop_odds_from_hist(hist2: np.ndarray, team_names: Sequence[str], usage: Sequence[TeamUsage],
                               team: str, player: str, stat: str, line: float) -> Dict[str, object]:
    """`edge_finder.player_prop_odds` (edge_finder.py:168-231) from the per-player histograms of one matchup
    (`[2][n_slots][PH_BINS]`, merged over ranks): same keys and values as from the per-game rows."""
    from .outputs import prob_to_american
    col = _STAT_ALIASES.get(stat, stat)
    if col not in _STAT_FIELD:
        raise ValueError(f"Stat '{stat}' (mapped to '{col}') not present in the player histograms.")
    role, fld = _STAT_FIELD[col]
    t = [i for i, nm in enumerate(team_names) if nm.lower() == team.lower()]
    hit = [s for s, (r, nm) in enumerate(usage[t[0]].slots) if r == role and nm.lower() == player.lower()] if t else []
    if not hit:
        raise ValueError(f"No rows found for {player} on {team}.")
    rec = np.asarray(hist2[t[0], hit[0]], dtype=np.int64)
    if fld == 0:
        counts = rec[:PH_YDS_BINS]
        values = (np.arange(PH_YDS_BINS, dtype=np.float64) - PH_YDS_OFFSET) / 10.0
    else:
        lo = PH_YDS_BINS + (fld - 1) * PH_CNT_BINS
        counts = rec[lo:lo + PH_CNT_BINS]
        values = np.arange(PH_CNT_BINS, dtype=np.float64)
    n = int(counts.sum())
    if n == 0:
        raise ValueError(f"No rows found for {player} on {team}.")
    p_over = float(counts[values > line].sum() / n)
    p_under = float(counts[values < line].sum() / n)
    p_push = float(counts[np.isclose(values, line, atol=1e-9)].sum() / n)
    ev = lambda p: p * (100.0 * (100.0 / 110.0)) - (1.0 - p) * 100.0
    implied = 110.0 / 210.0
    side, best_ev, edge = (("Over", ev(p_over), p_over - implied) if ev(p_over) >= ev(1.0 - p_over)
                           else ("Under", ev(1.0 - p_over), (1.0 - p_over) - implied))
    nz = counts > 0
    return {
        "team": team, "player": player, "role": ROLE_LABEL[role], "stat": col, "line": float(line),
        "samples": n, "p_over": round(p_over, 4), "p_under": round(p_under, 4),
        "push_rate": round(p_push, 4), "american_over": prob_to_american(p_over),
        "american_under": prob_to_american(p_under), "mean": float((values[nz] * counts[nz]).sum() / n),
        "median": _percentile_from_counts(values, counts, 50.0),
        "p75": _percentile_from_counts(values, counts, 75.0), "p90": _percentile_from_counts(values, counts, 90.0),
        "best_side": side, "edge": round(edge * 100, 2), "ev_per_$100": round(best_ev, 2),
    }


def scan_props_from_hist(hist2: np.ndarray, team_names: Sequence[str], usage: Sequence[TeamUsage],
                         prop_sheet_path: str, min_abs_edge_pct: float = 0.0) -> pd.DataFrame:
    """`edge_finder.scan_props_for_matchup` (edge_finder.py:340-390): every line of the prop sheet
    (`team, player, stat, yards`) that belongs to one of the two teams, priced from the histograms; rows the
    simulation has nothing for are skipped, the rest sorted by |edge| then EV."""
    cols = ["team", "player", "stat", "line", "best_side", "p_over", "p_under", "edge_pct", "ev_$100", "mean", "median", "samples"]
    if not os.path.exists(prop_sheet_path):
        return pd.DataFrame(columns=cols)
    props = pd.read_csv(prop_sheet_path)
    low = {t.lower() for t in team_names}
    keep = props[props["team"].astype(str).str.lower().isin(low)]
    rows = []
    for _, r in keep.iterrows():
        stat = _STAT_ALIASES.get(str(r["stat"]), str(r["stat"]))
        try:
            o = player_prop_odds_from_hist(hist2, team_names, usage, str(r["team"]), str(r["player"]), stat, float(r["yards"]))
        except Exception:
            continue
        rows.append({"team": r["team"], "player": r["player"], "stat": stat, "line": float(r["yards"]),
                     "best_side": o["best_side"], "p_over": o["p_over"], "p_under": o["p_under"], "edge_pct": o["edge"],
                     "ev_$100": o["ev_per_$100"], "mean": o["mean"], "median": o["median"], "samples": o["samples"]})
    if not rows:
        return pd.DataFrame(columns=cols)
    df = pd.DataFrame(rows)
    df["abs_edge"] = df["edge_pct"].abs()
    df = df.sort_values(["abs_edge", "ev_$100"], ascending=[False, False])
    return df[df["abs_edge"] >= min_abs_edge_pct].drop(columns=["abs_edge"])
