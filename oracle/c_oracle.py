"""ORACLE (test infrastructure only): ctypes front end of oracle/fmc_oracle.c.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfmc_oracle.so")

N_SLOTS = 16
MAX_ITERS = 360
TRACE_COLS = 8
SLOT = dict(U_call=0, U_comp=1, Z_yards=2, U_ex=3, U_boost=4, U_fin=5, U_s2=6, Z_int=7,
            U_go=8, U_fg=9, Z_gross=10, Z_ret=11, U_tb=12, U_p1=13, U_wr=14, U_yq=15)
NORMAL_SLOTS = (2, 7, 10, 11)
MODEL_IDS = {"pass_stage1": 0, "pass_stage2": 1, "pass_yards": 2, "run_yards": 3, "sack_yards": 4,
             "play_model": 5, "play_binary": 5, "run_fumble": 6}
COUNTER_NAMES = ("plays", "iters", "pass", "comp", "inc", "int", "sack", "run", "td", "fga", "fg", "punt", "go")
STAGE2_STANDIN = tuple(float(np.float32(x)) for x in (0.78, 0.05, 0.17))


class FoConfig(C.Structure):
    _fields_ = [
        ("sp", (C.c_double * 3) * 2),
        ("active", (C.c_int * 2) * 7),
        ("coach_col", C.c_int * 2),
        ("policy", C.c_int),
        ("play_temp", C.c_double),
        ("sampler", C.c_int),
        ("qy_noise", C.c_double),
        ("stage2_mode", C.c_int),
        ("standin", C.c_double * 3),
        ("pass_class", C.c_int),
    ]


MAX_USAGE = 32
PLAYER_FIELDS = 6
ROLE_INDEX = {"pass": 0, "rush": 1, "rec": 2}


class FoUsage(C.Structure):
    _fields_ = [("n", C.c_int), ("share", C.c_double * MAX_USAGE), ("slot", C.c_int * MAX_USAGE),
                ("col", (C.c_int * MAX_USAGE) * 7)]


class FoTeamUsage(C.Structure):
    _fields_ = [("role", FoUsage * 3)]


def make_usage(team_usages):
    """Two fast_monte_carlo_b200.usage.TeamUsage-like objects (role[r].share/.slot/.col) -> FoTeamUsage[2]."""
    arr = (FoTeamUsage * 2)()
    for t, tu in enumerate(team_usages):
        for rname, ri in ROLE_INDEX.items():
            ru = tu.role[rname]
            u = arr[t].role[ri]
            u.n = len(ru.names)
            for e in range(MAX_USAGE):
                u.slot[e] = -1
                for m in range(7):
                    u.col[m][e] = -1
            for e in range(u.n):
                u.share[e] = float(ru.share[e])
                u.slot[e] = int(ru.slot[e])
                for name, mid in MODEL_IDS.items():
                    if name in ru.col:
                        u.col[mid][e] = int(ru.col[name][e])
    return arr


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "fmc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libfmc_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.fo_ppnd16.restype = C.c_double
        _lib.fo_ppnd16.argtypes = [C.c_double]
        _lib.fo_pass_prob_v1.restype = C.c_double
        _lib.fo_pass_prob_v1.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        _lib.fo_go_for_it_prob.restype = C.c_double
        _lib.fo_go_for_it_prob.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        _lib.fo_field_goal_prob.restype = C.c_double
        _lib.fo_field_goal_prob.argtypes = [C.c_double]
        _lib.fo_modifiers.restype = None
        _lib.fo_modifiers.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def load_models(ms, play: str = "play_model") -> None:
    """`play`: which forest occupies the play-model slot: "play_model" (play_model.xgb) or "play_binary"
    (play_model.json)."""
    L = lib()
    for name, f in ms.forests.items():
        if name in ("play_model", "play_binary") and name != play:
            continue
        mid = MODEL_IDS[name]
        base = np.ascontiguousarray(f.base_margin, dtype=np.float64)
        arrs = dict(feat=np.ascontiguousarray(f.feat, np.int32), thr=np.ascontiguousarray(f.thr, np.float32),
                    left=np.ascontiguousarray(f.left, np.int32), right=np.ascontiguousarray(f.right, np.int32),
                    dl=np.ascontiguousarray(f.default_left, np.uint8), value=np.ascontiguousarray(f.value, np.float64),
                    root=np.ascontiguousarray(f.tree_root, np.int32), out=np.ascontiguousarray(f.tree_out, np.int32))
        rc = L.fo_load_forest(
            mid, int(f.kind), int(f.n_outputs), int(f.n_features), int(f.num_base), int(f.n_num),
            int(bool(f.zero_is_missing)), _p(base, C.c_double), C.c_double(float(f.scale)), int(f.n_nodes),
            _p(arrs["feat"], C.c_int), _p(arrs["thr"], C.c_float), _p(arrs["left"], C.c_int),
            _p(arrs["right"], C.c_int), _p(arrs["dl"], C.c_ubyte), _p(arrs["value"], C.c_double),
            int(f.n_trees), _p(arrs["root"], C.c_int), _p(arrs["out"], C.c_int))
        assert rc == 0, (name, rc)
        if f.scaler_cols is not None:
            cols = np.ascontiguousarray(f.scaler_cols, np.int32)
            mean = np.ascontiguousarray(f.scaler_mean, np.float64)
            sc = np.ascontiguousarray(f.scaler_scale, np.float64)
            L.fo_set_scaler(mid, int(cols.shape[0]), _p(cols, C.c_int), _p(mean, C.c_double), _p(sc, C.c_double))


def predict(name: str, num17: np.ndarray, active: np.ndarray, n_outputs: int,
            tree_begin: int = 0, tree_end: int = -1) -> np.ndarray:
    L = lib()
    num = np.zeros((num17.shape[0], 17), dtype=np.float64)
    num[:, :num17.shape[1]] = num17
    act = np.ascontiguousarray(np.asarray(active, dtype=np.int32).reshape(num.shape[0], 2))
    out = np.zeros((num.shape[0], n_outputs), dtype=np.float64)
    rc = L.fo_predict(MODEL_IDS[name], C.c_long(num.shape[0]), _p(num, C.c_double), _p(act, C.c_int),
                      int(tree_begin), int(tree_end), _p(out, C.c_double))
    assert rc == 0, rc
    return out


def make_config(ms, spA, spB, *, policy="heuristic", coach_cols=(-1, -1), play_temp=1.0, sampler="normal",
                qy_noise=0.5, stage2="standin", player="Unknown", pass_class=None) -> FoConfig:
    cfg = FoConfig()
    for t, sp in enumerate((spA, spB)):
        for k in range(3):
            cfg.sp[t][k] = float(sp[k])
    for name, mid in MODEL_IDS.items():
        cols = [-1, -1]
        if name in ms.forests:
            for gi, g in enumerate(ms[name].groups[:2]):
                if g.name != "coach":
                    cols[gi] = g.column_of(player)
        cfg.active[mid][0], cfg.active[mid][1] = cols
    cfg.coach_col[0], cfg.coach_col[1] = int(coach_cols[0]), int(coach_cols[1])
    cfg.policy = {"heuristic": 0, "play_model": 1, "play_json": 1}[policy]
    if pass_class is None:
        pass_class = ms["play_binary"].extra["pass_class"] if policy == "play_json" else 1
    cfg.pass_class = int(pass_class)
    if policy == "play_json":
        coach_cols = (-1, -1)
    cfg.play_temp = float(play_temp)
    cfg.sampler = {"normal": 0, "quantile_interp": 1}[sampler]
    cfg.qy_noise = float(qy_noise)
    cfg.stage2_mode = {"standin": 0, "booster": 1}[stage2]
    for k in range(3):
        cfg.standin[k] = STAGE2_STANDIN[k]
    return cfg


def simulate(cfg: FoConfig, n: int, *, game0: int = 0, matchup: int = 0, stream: np.ndarray | None = None,
             seed: int = 0, trace: bool = False, threads: int = 0, usage=None, n_slots: int = 0):
    """Returns dict(scores[n,2] (team A, team B), iters[n], trace[n,360,8]|None, counters{...});
    with `usage` (make_usage) also players[n,2,n_slots,6] = yds, att|tgt, comp|rec, td, INT, sacks."""
    L = lib()
    scores = np.zeros((n, 2), dtype=np.int32)
    iters = np.zeros(n, dtype=np.int32)
    tr = np.full((n, MAX_ITERS, TRACE_COLS), np.nan, dtype=np.float64) if trace else None
    counters = np.zeros(16, dtype=np.int64)
    if stream is not None:
        stream = np.ascontiguousarray(stream, dtype=np.float64)
        assert stream.shape == (n, MAX_ITERS, N_SLOTS), stream.shape
    tail = (C.c_long(n), C.c_long(game0), int(matchup), 0 if stream is not None else 1,
            _p(stream, C.c_double) if stream is not None else None, C.c_uint64(seed),
            _p(scores, C.c_int), _p(iters, C.c_int), _p(tr, C.c_double) if trace else None,
            _p(counters, C.c_long), int(threads))
    players = None
    if usage is None:
        rc = L.fo_simulate(C.byref(cfg), *tail)
    else:
        players = np.zeros((n, 2, max(n_slots, 1), PLAYER_FIELDS), dtype=np.float64)
        rc = L.fo_simulate_players(C.byref(cfg), usage, int(max(n_slots, 1)), _p(players, C.c_double), *tail)
        players = players[:, :, :n_slots, :]
    assert rc == 0, rc
    return dict(scores=scores, iters=iters, trace=tr, players=players,
                counters={k: int(counters[i]) for i, k in enumerate(COUNTER_NAMES)})


def play_pass_prob(cfg: FoConfig, off: int, down: int, dist: float, ytg: float, sd: int, sec: int) -> float:
    L = lib()
    L.fo_play_pass_prob.restype = C.c_double
    L.fo_play_pass_prob.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    return float(L.fo_play_pass_prob(C.byref(cfg), int(off), int(down), float(dist), float(ytg), int(sd), int(sec)))


def philox(ctr, key):
    L = lib()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    L.fo_philox(c, k, o)
    return tuple(int(x) for x in o)


def make_stream(n_games: int, seed: int) -> np.ndarray:
    """The injected-draw tensor of BASELINE config 2: uniform slots U[0,1), normal slots N(0,1)."""
    rng = np.random.default_rng(seed)
    s = rng.random((n_games, MAX_ITERS, N_SLOTS))
    z = rng.standard_normal((n_games, MAX_ITERS, len(NORMAL_SLOTS)))
    for j, sl in enumerate(NORMAL_SLOTS):
        s[:, :, sl] = z[:, :, j]
    return s


def modifiers(off_offense: float, def_defense: float, ytg: float, down: int) -> dict:
    out = (C.c_double * 6)()
    lib().fo_modifiers(float(off_offense), float(def_defense), float(ytg), int(down), out)
    return dict(matchup_bias=out[0], yardage_multiplier=out[1], mismatch_z=out[2], explosive_prob=out[3],
                rz_finish_prob_pass=out[4], rz_finish_prob_run=out[5])
