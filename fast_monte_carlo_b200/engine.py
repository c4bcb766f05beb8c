"""Host-side engine: owns one GPU context, the compiled forests and the matchup table.

Mirrors the reference's module-level state: the models loaded at import (FMC:641-668), the play
policy switch (FMC:46, 326-328) and the per-process worker contexts (`_init_pool`, FMC:1306-1319),
but as an object so that several engines (one per GPU / rank) can coexist.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import artifacts as art
from . import native

# FMC:55-61
HEAD_COACH_MAP = {
    "Kansas State": "Chris Klieman",
    "Iowa State": "Matt Campbell",
    "Kansas": "Lance Leipold",
    "Fresno State": "Matt Entz",
}

SIM_MODELS = ("pass_stage1", "pass_stage2", "pass_yards", "run_yards", "sack_yards", "play_model")


@dataclass
class MatchupSpec:
    """One team pair as the kernels see it."""
    team_a: str
    team_b: str
    sp_a: Tuple[float, float, float]     # RATING, OFFENSE, DEFENSE  (lookup_sp_flex FMC:1625-1644)
    sp_b: Tuple[float, float, float]
    games: int                           # total games of this matchup in the run (all ranks)
    game_begin: int = 0                  # this rank's slice [game_begin, game_end)
    game_end: int = 0
    out_offset: int = 0
    usage: Optional[tuple] = None        # (usage.TeamUsage of team A, of team B) or None = every name "Unknown"


class Engine:
    def __init__(self, models: Optional[art.ModelSet] = None, *, device: int = 0,
                 policy: str = "heuristic", sampler: str = "normal", stage2: str = "auto",
                 play_temp: Optional[float] = None, qy_noise: float = 0.5,
                 stage2_standin: Sequence[float] = (0.78, 0.05, 0.17), player: str = "Unknown",
                 memo: Optional[str] = None, memo_bytes: int = 0):
        """`memo`: "on" (default) | "off" | "persistent" -- the exact rank-keyed memo in front of the tree walk, the
        engine's counterpart of the reference's own memo caches (FMC:68-94); results never depend on it.  The
        environment variable FMC_MEMO overrides the default (the GPU test-suite is run under both settings)."""
        self.models = models if models is not None else art.load_default_models()
        self.ctx = native.Context(device)
        self.memo = memo if memo is not None else os.environ.get("FMC_MEMO", "on")
        self.ctx.set_memo(self.memo, max_bytes=memo_bytes,
                          max_trips=int(os.environ.get("FMC_MEMO_TRIPS", "0")),
                          break_parked=int(os.environ.get("FMC_MEMO_BREAK", "0")))
        self.player = player
        if policy == "play_json" and "play_binary" not in self.models:
            raise ValueError("policy='play_json' but the model set has no play_binary forest (play_model.json + "
                             "features.pkl + label_encoder.pkl in FMC_MODEL_DIR)")
        # the play-model slot holds ONE forest: play_model.json (policy 'play_json', what FMC:319-337 loads) or
        # play_model.xgb (policy 'play_model')
        play_name = "play_binary" if policy == "play_json" else "play_model"
        for name in SIM_MODELS[:-1] + (play_name, "run_fumble"):
            if name in self.models:
                f = self.models[name]
                mid = art.MODEL_IDS[name]
                self.ctx.load_forest(mid, f)
                cols = [-1, -1]
                if name not in ("play_model", "play_binary"):
                    for gi, g in enumerate(f.groups[:2]):
                        cols[gi] = g.column_of(player)
                self.ctx.set_active_columns(mid, cols[0], cols[1])
        if stage2 == "auto":
            stage2 = "booster" if "pass_stage2" in self.models else "standin"
        if stage2 == "booster" and "pass_stage2" not in self.models:
            raise ValueError("stage2='booster' but the model set has no pass_stage2 forest")
        if policy == "play_model" and "play_model" not in self.models:
            raise ValueError("policy='play_model' but the model set has no play_model forest")
        self.policy, self.sampler, self.stage2 = policy, sampler, stage2
        pass_class = 1
        if policy == "play_json":
            pass_class = int(self.models["play_binary"].extra["pass_class"])
            if play_temp is None:
                play_temp = float(self.models["play_binary"].extra.get("temperature", 1.0))   # calibration.json
        if play_temp is None:
            play_temp = 1.0
        self.play_temp = float(play_temp)
        self.ctx.set_params(policy={"heuristic": 0, "play_model": 1, "play_json": 1}[policy], pass_class=pass_class,
                            sampler={"normal": 0, "quantile_interp": 1}[sampler],
                            stage2_mode={"standin": 0, "booster": 1}[stage2],
                            play_temp=play_temp, qy_noise=qy_noise, stage2_standin=stage2_standin)
        self.matchups: List[MatchupSpec] = []

    # ------------------------------------------------------------------------------------------
    def coach_col(self, team: str) -> int:
        if self.policy == "play_json" or "play_model" not in self.models:
            return -1      # play_model.json: the category code the booster sees is 0 for every team (FMC:310-313)
        g = self.models["play_model"].group("coach")
        return g.column_of(HEAD_COACH_MAP.get(team)) if g is not None else -1

    def set_matchups(self, specs: Iterable[MatchupSpec]) -> None:
        self.matchups = list(specs)
        self.ctx.set_matchups([
            dict(sp=[list(m.sp_a), list(m.sp_b)], coach_col=(self.coach_col(m.team_a), self.coach_col(m.team_b)),
                 game_begin=m.game_begin, game_end=m.game_end, out_offset=m.out_offset)
            for m in self.matchups])
        # usage tables (FMC:228-249): all matchups or none; a slate mixing both gives the others the trivial table
        with_usage = [m for m in self.matchups if m.usage is not None and not (m.usage[0].trivial and m.usage[1].trivial)]
        if with_usage:
            from . import usage as _usage
            teams = []
            for m in self.matchups:
                if m.usage is None:
                    raise ValueError("player usage must be given for every matchup of a slate or for none")
                teams.append(m.usage)
            self.n_slots = max(len(tu.slots) for pair in teams for tu in pair)
            self.ctx.set_usage(teams, self.n_slots)
        else:
            self.n_slots = 0

    def simulate_host(self, seed: int, **kw) -> dict:
        return self.ctx.simulate_host(seed=seed, **kw)

    def predict(self, name: str, rows: np.ndarray, tree_begin: int = 0, tree_end: int = -1,
                coach: Optional[str] = None, names: Optional[Sequence[Sequence[Optional[str]]]] = None,
                hot_cols: Optional[np.ndarray] = None) -> np.ndarray:
        """Raw margins of one model on [n,17] numerics.  Without `names` / `hot_cols` every row has this engine's
        `player` ("Unknown") in its name columns.  `names`: per row the values of the model's categorical inputs in the
        order of `forest.groups` (passer_name, target_name | rusher_name | coach), exactly what a DataFrame row of the
        reference carries (FMC:744, 756, 784-809); names that are not categories light nothing.  `hot_cols` gives the
        one-hot columns directly (int32 [n, 2], -1 = none)."""
        f = self.models[name]
        if names is not None:
            groups = f.groups[:2]
            hot_cols = np.full((len(names), 2), -1, dtype=np.int32)
            for i, row in enumerate(names):
                for gi, g in enumerate(groups):
                    if gi < len(row):
                        hot_cols[i, gi] = g.column_of(row[gi])
        if hot_cols is not None:
            return self.ctx.tree_predict_cols_host(art.MODEL_IDS[name], rows, hot_cols, f.n_outputs, tree_begin, tree_end)
        cc = -1
        if name == "play_model" and coach is not None:
            cc = f.group("coach").column_of(coach)
        return self.ctx.tree_predict_host(art.MODEL_IDS[name], rows, f.n_outputs, tree_begin, tree_end, cc)

    def close(self) -> None:
        self.ctx.close()
