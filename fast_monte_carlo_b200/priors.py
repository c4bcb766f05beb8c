"""SP+ priors: the team-rating inputs of the simulator.

Same behaviour as the reference's loader (fast_monte_carlo_cfb.py:1573-1659): two accepted CSV
schemas, punctuation/case-insensitive team lookup with the same three fallbacks, and the same
`TeamContext` fields.  Stays on the host: it runs once per matchup.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import pandas as pd

_SCHEMA_A = ("team", "RATING", "OFFENSE", "DEFENSE")
_SCHEMA_B = ("Current SP+", "Past SP+", "Rating", "Offense Rating", "Defense Rating")
_TABLES: Dict[str, pd.DataFrame] = {}


def _norm_team(s) -> str:
    """Lower-case and drop everything that is not a letter or digit (FMC:1573-1574)."""
    return re.sub(r"[^a-z0-9]+", "", str(s).lower())


def packaged_priors_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sp_2025_week1.csv")


def load_sp_flex(sp_path: str) -> pd.DataFrame:
    """CSV -> frame with columns team, RATING, OFFENSE, DEFENSE, norm_team (FMC:1576-1623).

    Schema A already has those columns.  Schema B (the 2025 file) lists every team under both its
    current and its past name; both spellings become rows, first occurrence wins.
    """
    if sp_path in _TABLES:
        return _TABLES[sp_path]
    raw = pd.read_csv(sp_path)
    have = set(raw.columns)
    if have.issuperset(_SCHEMA_A):
        sp = raw[list(_SCHEMA_A)].copy()
    elif have.issuperset(_SCHEMA_B):
        ratings = ["Rating", "Offense Rating", "Defense Rating"]
        halves = []
        for name_col in ("Current SP+", "Past SP+"):
            part = raw[[name_col] + ratings].copy()
            part.columns = list(_SCHEMA_A)
            halves.append(part)
        sp = pd.concat(halves, ignore_index=True)
        sp = sp[sp["team"].notna()].copy()
        sp["team"] = sp["team"].astype(str).str.strip()
        sp = sp.drop_duplicates(subset=["team"], keep="first")
    else:
        raise ValueError(
            f"Unrecognized SP+ schema in {sp_path}. Expected columns either "
            f"[team,RATING,OFFENSE,DEFENSE,...] or "
            f"['Current SP+','Past SP+','Rating','Offense Rating','Defense Rating']")
    sp["team"] = sp["team"].astype(str).str.strip()
    for c in ("RATING", "OFFENSE", "DEFENSE"):
        sp[c] = sp[c].astype(float)
    sp["norm_team"] = sp["team"].map(_norm_team)
    _TABLES[sp_path] = sp
    return sp


def lookup_sp_flex(team: str, sp_df: pd.DataFrame) -> Tuple[float, float, float]:
    """(RATING, OFFENSE, DEFENSE): normalised match, then lower-case match, then substring (FMC:1625-1644)."""
    hit = sp_df[sp_df["norm_team"] == _norm_team(team)]
    if hit.empty:
        hit = sp_df[sp_df["team"].str.lower() == team.lower()]
    if hit.empty:
        loose = sp_df[sp_df["team"].str.lower().str.contains(team.lower(), regex=False)]
        if not loose.empty:
            hit = loose.iloc[:1]
    if hit.empty:
        raise ValueError(f"Team '{team}' not found in provided SP+ table.")
    r = hit.iloc[0]
    return float(r["RATING"]), float(r["OFFENSE"]), float(r["DEFENSE"])


def _unknown_share(col: str) -> pd.DataFrame:
    return pd.DataFrame({col: ["Unknown"], "share": [1.0]})


@dataclass
class TeamContext:
    """Field-compatible with the reference's TeamContext (FMC:255-271)."""
    name: str
    year: int
    week: int
    sp_rating: float
    sp_offense: float
    sp_defense: float
    qb_share: Optional[pd.DataFrame] = None
    rush_share: Optional[pd.DataFrame] = None
    target_share: Optional[pd.DataFrame] = None
    track_pass: Optional[set] = None
    track_rush: Optional[set] = None
    track_rec: Optional[set] = None

    @property
    def sp(self) -> Tuple[float, float, float]:
        return (self.sp_rating, self.sp_offense, self.sp_defense)


def build_team_context_from_sp_flex(team: str, year: int, week: int, sp_df: pd.DataFrame, *,
                                    focus: Optional[dict] = None, usage_dir: str = ".") -> TeamContext:
    """FMC:1646-1659.  Usage comes from the focus sheet (`usage.build_focus_usage_tables`), else from
    `usage_*_share.csv` under `usage_dir`, else every passer / rusher / target is "Unknown" with share
    1.0 and nothing is tracked (FMC:228-249) -- the only configuration the shipped reference can reach."""
    from .usage import usage_for_team
    rating, off, de = lookup_sp_flex(team, sp_df)
    qb, ru, tg, tp, tr, trec = usage_for_team(team, year, focus, usage_dir)
    return TeamContext(name=team, year=year, week=week, sp_rating=rating, sp_offense=off, sp_defense=de,
                       qb_share=qb, rush_share=ru, target_share=tg,
                       track_pass=tp, track_rush=tr, track_rec=trec)


def csv_base_from(team_a: str, team_b: str, week: int, ext: str = ".csv") -> str:
    """`<norm(a)>_<norm(b)>_wk<week>_sims<ext>` (FMC:1717-1722)."""
    return f"{_norm_team(team_a)}_{_norm_team(team_b)}_wk{int(week)}_sims{ext}"
