"""Drop-in Python entry points of the reference (fast_monte_carlo_cfb.py:1467-1521, 1661-1722),
executed by the B200 engine.

    simulate_upcoming_matchup(teamA, teamB, *, year, week, sp_path, n, show_progress,
                              collect_players, save_csv, processes)
        -> (sims_df, players_df, summary, A, B, meta)
    simulate_matchup(teamA_ctx, teamB_ctx, n, seed, show_progress, collect_players, players_csv,
                     processes) -> (sims_df, players_df | None)

Same argument meaning, return shapes, file names (`scores_<save_csv>`, `players_<save_csv>`) and
error behaviour (ValueError for an unknown team / schema).  `n` counts PAIRS of games (A receives,
then B receives); `processes` and `show_progress` are accepted and ignored (the GPU replaces the
process pool).  Differences, all documented in DESIGN.md: a given `seed` makes the run
reproducible game by game (counter-based Philox keyed by (seed, game id)) instead of re-seeding
NumPy before every game.

Players: when a team has usage data -- the focus sheet `2025_week1_players.csv`
(FMC:508-605) or `usage_*_share.csv` (FMC:487-505), looked up in the working directory exactly like
the reference, or passed with `focus_csv=` / `usage_dir=` -- passer, target and rusher are sampled
per play on the GPU, feed the models' one-hot columns, and the focus names get per-game box lines:
`players_df` has the reference's PLAYER_COLS rows (`flatten_player_box_rows`, FMC:1266-1299).  Without
usage data every name is "Unknown" and `players_df` is empty, as in the shipped reference.
"""
from __future__ import annotations

import os
import threading
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import outputs
from . import usage as _usage
from .engine import Engine, MatchupSpec
from .priors import (TeamContext, build_team_context_from_sp_flex, csv_base_from, load_sp_flex,
                     lookup_sp_flex, packaged_priors_path)

# FMC:1259-1264
PLAYER_COLS = [
    "sim", "start", "team", "opp", "player", "role",
    "pass_att", "pass_comp", "pass_yds", "pass_td", "INT", "sacks",
    "rush_att", "rush_yds", "rush_td",
    "rec", "tgt", "rec_yds", "rec_td",
]

_ENGINES: Dict[tuple, Engine] = {}
# The engine of a device is shared process-wide (like the reference's module-level models) and holds the matchup list
# of the call in progress: calls from several threads are serialised.
_ENGINE_LOCK = threading.RLock()
# joint score histogram [2, 128, 128] + event counters of the most recent simulate_matchup call
LAST_RUN: Dict[str, object] = {}


def get_engine(device: Optional[int] = None, **kw) -> Engine:
    """Process-wide engine per (device, options) -- the analogue of the reference's module-level
    model objects.  Raises when the CUDA library or a GPU is missing: there is no CPU path."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    key = (device,) + tuple(sorted(kw.items()))
    if key not in _ENGINES:
        _ENGINES[key] = Engine(device=device, **kw)
    return _ENGINES[key]


class _Attached:
    """A value riding in DataFrame.attrs: compared by identity, so that pandas' own attrs comparisons (concat, astype)
    never evaluate an array for truth."""
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value

    def __eq__(self, other):
        return self is other

    def __hash__(self):
        return id(self)


def _fresh_seed() -> int:
    return int.from_bytes(os.urandom(8), "little")


def simulate_matchup(teamA: TeamContext, teamB: TeamContext, n: int = 100, seed: Optional[int] = None,
                     show_progress: bool = True, collect_players: bool = False,
                     players_csv: Optional[str] = None, processes: Optional[int] = None,
                     *, engine: Optional[Engine] = None,
                     game_range: Optional[Tuple[int, int]] = None) -> Tuple[pd.DataFrame, Optional[pd.DataFrame]]:
    """`game_range=(g0, g1)` (ours, for one process per GPU): simulate only games [g0, g1) of the 2n -- `shard_range`
    gives a rank its slice; the rows returned are those games, and the histograms of the ranks add up to the run's
    (`merge_histograms`).  A given seed gives the same games whatever the split."""
    eng = engine if engine is not None else get_engine()
    games = 2 * int(n)
    g0, g1 = (0, games) if game_range is None else (int(game_range[0]), int(game_range[1]))
    if not (0 <= g0 <= g1 <= max(games, 0)):
        raise ValueError("game_range must lie inside [0, 2n]")
    box = None
    use = (_usage.resolve_team(teamA, eng.models), _usage.resolve_team(teamB, eng.models))
    if g1 - g0 <= 0:
        sims_df = pd.DataFrame(columns=["team", "opp", "pts", "opp_pts"])
    else:
        # the name columns of the result depend on the names and the game parity only: built while the GPU plays
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as side:
            names_f = side.submit(outputs.name_columns, teamA.name, teamB.name, g1 - g0, g0)
            with _ENGINE_LOCK:
                # the sampled names feed the models whether or not the box is collected (FMC:1058-1081, 1203-1216)
                eng.set_matchups([MatchupSpec(teamA.name, teamB.name, teamA.sp, teamB.sp, games, g0, g1, 0, usage=use)])
                want_box = bool(collect_players) and eng.n_slots > 0
                res = eng.simulate_host(_fresh_seed() if seed is None else int(seed), want_scores=True, want_hist=True,
                                        want_players=want_box)
            names = names_f.result()
        box = res.get("players")
        sims_df = outputs.sims_frame(teamA.name, teamB.name, res["scores"], first_game=g0, names=names)
        sims_df.attrs["counters"] = dict(res["counters"])
        sims_df.attrs["hist"] = _Attached(res["hist"][0])
        LAST_RUN.clear()
        LAST_RUN.update(hist=res["hist"][0], counters=dict(res["counters"]), teams=(teamA.name, teamB.name),
                        scores=res["scores"], player_box=box, usage=use, game_range=(g0, g1))
    players_df = None
    if collect_players:
        players_df = (_usage.player_rows(box, 0, (teamA.name, teamB.name), use) if box is not None
                      else pd.DataFrame(columns=PLAYER_COLS))
        if players_csv:
            if players_csv.lower().endswith(".parquet"):
                players_df.to_parquet(players_csv, index=False)
            else:
                players_df.to_csv(players_csv, index=False)
    return sims_df, players_df


def simulate_upcoming_matchup(teamA: str, teamB: str, *, year: int = 2025, week: int = 1,
                              sp_path: str = "Pregame_SPPlus2025_1.csv", n: int = 1000,
                              show_progress: bool = True, collect_players: bool = True,
                              save_csv: Optional[str] = None, processes: Optional[int] = None,
                              seed: Optional[int] = None, engine: Optional[Engine] = None,
                              focus_csv: Optional[str] = None, usage_dir: str = "."):
    sp_df = load_sp_flex(sp_path)
    # FMC:605 builds the focus tables at import from `2025_week1_players.csv` in the working directory
    focus = _usage.build_focus_usage_tables(focus_csv if focus_csv is not None else _usage.FOCUS_PLAYERS_CSV)
    A = build_team_context_from_sp_flex(teamA, year, week, sp_df, focus=focus, usage_dir=usage_dir)
    B = build_team_context_from_sp_flex(teamB, year, week, sp_df, focus=focus, usage_dir=usage_dir)

    t0 = time.perf_counter()
    sims_df, players_df = simulate_matchup(A, B, n=n, seed=seed, show_progress=show_progress,
                                           collect_players=collect_players, processes=processes, engine=engine)
    t1 = time.perf_counter()
    # FMC:1681-1687.  The group-by over the 2n rows and the joint score histogram the kernel accumulated hold the same
    # information; the histogram gives the five statistics in microseconds (tests check the two agree)
    c = sims_df.attrs.get("counters")
    h = sims_df.attrs.get("hist")
    h = h.value if isinstance(h, _Attached) else None
    if h is not None and c and c.get("hist_overflow", 1) == 0 and min(int(h[0].sum()), int(h[1].sum())) > 1:
        summary = outputs.summary_from_hist(h, A.name, B.name)
    else:
        summary = outputs.summary_frame(sims_df)

    write_time = 0.0
    if save_csv:
        tw = time.perf_counter()
        try:
            if save_csv.lower().endswith(".parquet"):
                raw = sims_df.attrs.get("scores_raw")
                if raw is None:
                    raw = LAST_RUN.get("scores") if LAST_RUN.get("teams") == (A.name, B.name) else None
                if raw is not None and len(raw) == len(sims_df):   # chunked, dictionary-encoded writer straight from the score array
                    outputs.write_scores_table(f"scores_{save_csv}", A.name, B.name, raw)
                else:
                    sims_df.to_parquet(f"scores_{save_csv}", index=False)
                if players_df is not None:
                    players_df.to_parquet(f"players_{save_csv}", index=False)
            else:
                sims_df.to_csv(f"scores_{save_csv}", index=False)
                if players_df is not None:
                    players_df.to_csv(f"players_{save_csv}", index=False)
        except Exception:
            sims_df.to_csv(f"scores_{save_csv}.csv", index=False)
            if players_df is not None:
                players_df.to_csv(f"players_{save_csv}.csv", index=False)
        write_time = time.perf_counter() - tw

    sim_time = t1 - t0
    meta = {"sim_time_sec": sim_time, "io_time_sec": write_time, "total_time_sec": sim_time + write_time, "sims": n}
    c = sims_df.attrs.get("counters")
    if c:
        meta["plays"] = c["plays"]
        meta["games"] = c["games"]
    return sims_df, players_df, summary, A, B, meta


# ---------------------------------------------------------------------------------------------
# slates (BASELINE configs 4 and 5): many matchups, games sharded over ranks, one histogram merge
# ---------------------------------------------------------------------------------------------
def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of [0, total) owned by `rank`; slices tile the range exactly."""
    return (total * rank) // world, (total * (rank + 1)) // world


def slate_specs(pairs: Sequence[Tuple[str, str]], games_per_matchup: int, sp_df: pd.DataFrame,
                rank: int = 0, world: int = 1, shard: str = "games") -> List[MatchupSpec]:
    """The matchup list of one rank.  Results do not depend on `world` or `shard`: the Philox key is (seed, matchup,
    game id) and the ranks' integer histograms add up (`merge_histograms`).
      shard="games":    every rank simulates its contiguous game-id slice of every matchup (FMC:1321-1328 treats a pair
                        of games as the unit of work);
      shard="matchups": rank r simulates all games of matchups r, r + world, ... and none of the others (their ranges are
                        empty) -- a matchup's node tables and its memo entries then live on one GPU only, which keeps the
                        memo's hit rate at that of an unsharded matchup."""
    if shard not in ("games", "matchups"):
        raise ValueError("shard must be 'games' or 'matchups'")
    specs = []
    off = 0
    for i, (a, b) in enumerate(pairs):
        spa, spb = lookup_sp_flex(a, sp_df), lookup_sp_flex(b, sp_df)
        if shard == "games":
            g0, g1 = shard_range(int(games_per_matchup), rank, world)
        else:
            g0, g1 = (0, int(games_per_matchup)) if i % world == rank else (0, 0)
        specs.append(MatchupSpec(a, b, spa, spb, int(games_per_matchup), g0, g1, off))
        off += g1 - g0
    return specs


def merge_histograms(hist, counters=None):
    """The single exchange step of the multi-GPU path: all-reduce(sum) of the integer histograms
    (+ counters) over the default torch.distributed group (NCCL over NVLink on GPUs, gloo on CPU).
    `hist` / `counters` are int64 torch tensors and are reduced in place.  No-op without a group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        if counters is not None:
            dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return hist, counters


def gather_player_box(local_rec, total_games: int):
    """The exchange step of the player path across ranks: every rank holds the per-game box of its contiguous
    game-id slice (`shard_range`), fmc_player_rec[games_local][2][n_slots]; one all-gather over the default
    torch.distributed group (NCCL for a CUDA tensor, gloo for a CPU tensor / NumPy array) returns the whole run's
    box in game order on every rank.  `local_rec` may be a NumPy structured array (native.PLAYER_REC), a
    native.PlayerBox or an int64 torch tensor [games_local, 2, n_slots, 2].  No-op without a process group.
    Memory: 32 bytes x n_slots x total_games on every rank -- meant for one matchup, not for a slate."""
    import torch
    import torch.distributed as dist
    from .native import PLAYER_REC, PlayerBox
    if isinstance(local_rec, PlayerBox):
        local_rec = local_rec.rec
    as_numpy = isinstance(local_rec, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local_rec).view(np.int64).reshape(local_rec.shape + (2,))) if as_numpy else local_rec
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return PlayerBox(local_rec) if as_numpy else t
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [shard_range(int(total_games), r, world) for r in range(world)]
    assert t.shape[0] == sizes[rank][1] - sizes[rank][0], "local box does not match this rank's game slice"
    longest = max(b - a for a, b in sizes)
    pad = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    whole = torch.cat([parts[r][:sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
    if as_numpy:
        return PlayerBox(whole.cpu().numpy().view(PLAYER_REC).reshape(whole.shape[:-1]))
    return whole


def simulate_slate(pairs: Sequence[Tuple[str, str]], n: int = 1000, *, sp_path: Optional[str] = None,
                   year: int = 2025, week: int = 1, seed: Optional[int] = None,
                   engine: Optional[Engine] = None, markets: Optional[Dict[Tuple[str, str], Dict[str, float]]] = None,
                   focus_csv: Optional[str] = None, usage_dir: str = "."):
    """A whole slate in one launch (BASELINE configs 3-4): every (teamA, teamB) of `pairs` is simulated for
    `n` PAIRS of games (2n games, A receives / B receives alternately, exactly like `simulate_matchup`).

    Under torch.distributed (one process per GPU) every rank plays its contiguous game-id slice of every
    matchup and the integer histograms are merged with ONE all-reduce; the result does not depend on the
    number of ranks.  No per-game table is materialised: everything `edge_finder.py` derives from
    `scores_*` -- `summary` (FMC:1681-1687), moneyline (edge_finder.py:249-281) and, for pairs listed in
    `markets` ({(A, B): {"spread": s, "total": t}}), spread / total odds (edge_finder.py:283-336) -- comes
    from the joint score histogram.  Returns {(A, B): {"hist", "summary", "moneyline", "markets", "games"}},
    plus the key "_counters" with the event counters of the whole slate.

    Players: when a team of the slate has usage data (`focus_csv`, default `2025_week1_players.csv` in the working
    directory, or `usage_*_share.csv` under `usage_dir`; FMC:228-249) the slate runs in player mode and every
    matchup also returns "player_hist" ([2][n_slots][PH_BINS], merged over ranks with the same all-reduce),
    "usage" and "props": `edge_finder.scan_props_for_matchup` (edge_finder.py:340-390) for the sheet's lines,
    computed from the histograms (`usage.player_prop_odds_from_hist`).
    """
    import torch
    import torch.distributed as dist
    eng = engine if engine is not None else get_engine()
    sp_df = load_sp_flex(sp_path if sp_path is not None else packaged_priors_path())
    sheet = focus_csv if focus_csv is not None else _usage.FOCUS_PLAYERS_CSV
    focus = _usage.build_focus_usage_tables(sheet)
    uses = []
    for a, b in pairs:                                    # same errors as the reference for unknown teams
        ta = build_team_context_from_sp_flex(a, year, week, sp_df, focus=focus, usage_dir=usage_dir)
        tb = build_team_context_from_sp_flex(b, year, week, sp_df, focus=focus, usage_dir=usage_dir)
        uses.append((_usage.resolve_team(ta, eng.models), _usage.resolve_team(tb, eng.models)))
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    specs = slate_specs(pairs, 2 * int(n), sp_df, rank, world)
    for sp_, u in zip(specs, uses):
        sp_.usage = u
    eng.set_matchups(specs)
    with_players = eng.ctx.has_usage and eng.n_slots > 0
    dev = torch.device("cuda", eng.ctx.device)
    hist = torch.zeros((len(specs), 2, outputs.HIST_BINS, outputs.HIST_BINS), dtype=torch.int32, device=dev)
    counters = torch.zeros(len(native_counter_names()), dtype=torch.int64, device=dev)
    padded = torch.zeros(32, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev)
    phist = (torch.zeros((len(specs), 2, eng.n_slots, _usage.PH_BINS), dtype=torch.int32, device=dev)
             if with_players else None)
    eng.ctx.simulate_device(seed=_fresh_seed() if seed is None else int(seed), hist=hist.data_ptr(),
                            counters=padded.data_ptr(), cuda_stream=st.cuda_stream,
                            player_hist=phist.data_ptr() if phist is not None else 0)
    hist64 = hist.to(torch.int64)
    counters.copy_(padded[:counters.numel()])
    merge_histograms(hist64, counters)
    ph = None
    if phist is not None:
        ph64 = phist.to(torch.int64)
        merge_histograms(ph64)
        ph = ph64.cpu().numpy()
    h = hist64.cpu().numpy()
    c = counters.cpu().numpy()
    out: Dict[object, object] = {"_counters": {k: int(c[i]) for i, k in enumerate(native_counter_names())}}
    for m, (a, b) in enumerate(pairs):
        entry = {"hist": h[m], "games": int(h[m].sum()), "summary": outputs.summary_from_hist(h[m], a, b),
                 "moneyline": outputs.moneyline_from_hist(h[m], a, b), "markets": None}
        mk = (markets or {}).get((a, b))
        if mk:
            entry["markets"] = outputs.game_market_odds_from_hist(h[m], a, b, spread=mk.get("spread"), total=mk.get("total"))
        if ph is not None:
            entry["player_hist"] = ph[m]
            entry["usage"] = uses[m]
            entry["props"] = _usage.scan_props_from_hist(ph[m], (a, b), uses[m], sheet)
        out[(a, b)] = entry
    return out


def native_counter_names():
    from .native import COUNTER_NAMES
    return COUNTER_NAMES
