"""The multi-rank path on CPU: world_size-2 gloo.  Each rank plays its contiguous game slice of every
matchup (here with the C oracle standing in for the GPU), histograms are merged with ONE all-reduce,
and the result must equal the single-process histogram -- independent of the number of ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_monte_carlo_b200 import api, outputs


def test_shard_ranges_tile():
    for total in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            edges = [api.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, games, q, shard="games"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fast_monte_carlo_b200 import artifacts as art, priors
    from oracle import c_oracle as co
    ms = art.load_default_models()
    co.load_models(ms)
    sp = priors.load_sp_flex(priors.packaged_priors_path())
    pairs = [("Kansas State", "Iowa State"), ("UTSA", "Ohio State"), ("Iowa State", "UTSA")]
    specs = api.slate_specs(pairs, games, sp, rank, world, shard=shard)
    hist = torch.zeros((len(pairs), 2, outputs.HIST_BINS, outputs.HIST_BINS), dtype=torch.int64)
    counters = torch.zeros(4, dtype=torch.int64)
    for m, s in enumerate(specs):
        if s.game_end == s.game_begin:
            continue                     # shard="matchups": this rank does not own the matchup
        cfg = co.make_config(ms, s.sp_a, s.sp_b)
        r = co.simulate(cfg, s.game_end - s.game_begin, game0=s.game_begin, matchup=m, seed=99, threads=1)
        hist[m] += torch.from_numpy(outputs.histogram_from_scores(r["scores"], s.game_begin))
        counters[0] += s.game_end - s.game_begin
        counters[1] += r["counters"]["plays"]
    api.merge_histograms(hist, counters)
    if rank == 0:
        q.put((hist.numpy(), counters.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_histogram_merge_equals_single_process():
    games = 300
    ctx = mp.get_context("spawn")
    results = {}
    for world, shard in ((1, "games"), (2, "games"), ("2m", "matchups")):
        n = 2 if world == "2m" else world
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, n, port, games, q, shard)) for r in range(n)]
        for p in procs:
            p.start()
        results[world] = q.get()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
    h1, c1 = results[1]
    h2, c2 = results[2]
    h3, c3 = results["2m"]               # whole matchups per rank instead of game slices
    assert np.array_equal(h1, h2) and np.array_equal(c1, c2)
    assert np.array_equal(h1, h3) and np.array_equal(c1, c3)
    assert h1.sum() == 3 * games and c1[0] == 3 * games


def _players_worker(rank, world, port, games, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import GOLDEN
    from fast_monte_carlo_b200 import artifacts as art, priors, usage, native
    from oracle import c_oracle as co
    ms = art.load_default_models()
    co.load_models(ms)
    focus = usage.build_focus_usage_tables(os.path.join(GOLDEN, "players_focus.csv"))
    sp = priors.load_sp_flex(priors.packaged_priors_path())
    tcs = [priors.build_team_context_from_sp_flex(t, 2025, 1, sp, focus=focus, usage_dir=GOLDEN)
           for t in ("Kansas State", "Iowa State")]
    us = [usage.resolve_team(tc, ms) for tc in tcs]
    n_slots = max(len(u.slots) for u in us)
    g0, g1 = api.shard_range(games, rank, world)
    r = co.simulate(co.make_config(ms, tcs[0].sp, tcs[1].sp), g1 - g0, game0=g0, seed=4, threads=1,
                    usage=co.make_usage(us), n_slots=n_slots)
    # pack the oracle's dense box into the kernel's record layout (fmc_player_rec) -- the rank-local product output
    rec = np.zeros((g1 - g0, 2, n_slots), dtype=native.PLAYER_REC)
    rec["yds"] = r["players"][..., 0]
    for k in range(5):
        rec["counts"] |= r["players"][..., 1 + k].astype(np.uint64) << np.uint64(10 * k)
    whole = api.gather_player_box(rec, games)
    assert whole.shape[0] == games
    odds = usage.player_prop_odds_from_box(whole, ("Kansas State", "Iowa State"), us, "Iowa State", "Levi Williams",
                                           "pass_yards", 150.5)
    if rank == 0:
        q.put((whole.dense(), odds))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_player_box_gather_equals_single_process():
    """Ranks play contiguous game-id slices; one all-gather of the per-game boxes gives the single-process box
    (odd game count: the slices differ in length), hence the same prop odds."""
    games = 301
    ctx = mp.get_context("spawn")
    results = {}
    for world in (1, 2):
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_players_worker, args=(r, world, port, games, q)) for r in range(world)]
        for p in procs:
            p.start()
        results[world] = q.get()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
    b1, o1 = results[1]
    b2, o2 = results[2]
    assert b1.shape[0] == games and np.array_equal(b1, b2) and o1 == o2 and o1["samples"] > 250
