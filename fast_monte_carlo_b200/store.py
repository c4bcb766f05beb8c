"""Result bundles in the layout of the reference's `sim_store.py` (sim_store.py:6-26): a run directory
with `games.parquet` (sim_id, team, opp, pts, opp_pts, margin, total, seed), `players.parquet` and
`meta.json`, plus the sha256 run signature.  Written from the engine's per-game score array, so a
10 M-game run is stored without building a pandas object frame."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import pandas as pd

GAMES_COLUMNS = ["sim_id", "team", "opp", "pts", "opp_pts", "margin", "total", "seed"]


def make_signature(meta: dict) -> str:
    """sha256 of the canonical JSON of `meta` (sim_store.py:6-8)."""
    s = json.dumps(meta, sort_keys=True, separators=(",", ":"))
    return hashlib.sha256(s.encode()).hexdigest()


def save_sim_bundle(run_dir: str, team_a: str, team_b: str, scores: np.ndarray, meta: dict,
                    players_df: Optional[pd.DataFrame] = None, seed: int = 0, first_game: int = 0,
                    chunk_rows: int = 4_000_000) -> str:
    """games.parquet / players.parquet / meta.json under `run_dir` (sim_store.py:10-19).  `scores` is the
    engine's [games, 2] (points A, points B) array; sim_id = game id; returns the run signature."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from .api import PLAYER_COLS
    p = Path(run_dir)
    p.mkdir(parents=True, exist_ok=True)
    names = pa.array([team_a, team_b], type=pa.string())
    dict_t = pa.dictionary(pa.int8(), pa.string())
    schema = pa.schema([("sim_id", pa.int64()), ("team", dict_t), ("opp", dict_t), ("pts", pa.int64()),
                        ("opp_pts", pa.int64()), ("margin", pa.int64()), ("total", pa.int64()), ("seed", pa.int64())])
    n = int(scores.shape[0])
    with pq.ParquetWriter(p / "games.parquet", schema, compression="zstd") as w:
        for lo in range(0, max(n, 1), chunk_rows):
            sc = scores[lo:lo + chunk_rows]
            g = np.arange(first_game + lo, first_game + lo + sc.shape[0], dtype=np.int64)
            b_first = (g & 1).astype(np.int8)
            pts = np.where(b_first == 0, sc[:, 0], sc[:, 1]).astype(np.int64)
            opp = np.where(b_first == 0, sc[:, 1], sc[:, 0]).astype(np.int64)
            w.write_table(pa.table({
                "sim_id": pa.array(g),
                "team": pa.DictionaryArray.from_arrays(pa.array(b_first), names),
                "opp": pa.DictionaryArray.from_arrays(pa.array((1 - b_first).astype(np.int8)), names),
                "pts": pa.array(pts), "opp_pts": pa.array(opp), "margin": pa.array(pts - opp),
                "total": pa.array(pts + opp), "seed": pa.array(np.full(sc.shape[0], int(seed) & 0x7FFFFFFFFFFFFFFF, np.int64)),
            }, schema=schema))
    if players_df is None:
        players_df = pd.DataFrame(columns=PLAYER_COLS)
    players_df.to_parquet(p / "players.parquet", index=False)
    meta = dict(meta)
    meta.setdefault("signature", make_signature({k: v for k, v in meta.items() if k != "signature"}))
    (p / "meta.json").write_text(json.dumps(meta, indent=2))
    return meta["signature"]


def load_sim_bundle(run_dir: str) -> Tuple[pd.DataFrame, pd.DataFrame, dict]:
    """(games, players, meta) exactly as sim_store.load_sim_bundle (sim_store.py:21-26)."""
    p = Path(run_dir)
    games = pd.read_parquet(p / "games.parquet")
    players = pd.read_parquet(p / "players.parquet")
    meta = json.loads((p / "meta.json").read_text())
    return games, players, meta
