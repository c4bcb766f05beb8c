"""Multi-GPU check of the player path (run under torchrun, one rank per GPU):
api.simulate_slate in player mode -- every rank plays its game-id slice of every matchup, score histograms and
per-player histograms are merged by NCCL all-reduces -- must equal rank 0 playing the whole slate alone.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/slate_players_n2.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from fast_monte_carlo_b200 import api, priors, usage
from fast_monte_carlo_b200.engine import MatchupSpec

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
gold = os.path.join(ROOT, "tests", "golden")
sheet = os.path.join(gold, "players_focus.csv")
pairs = [("Kansas State", "Iowa State"), ("Ohio State", "Kansas State"), ("UTSA", "Iowa State")]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000          # pairs of games per matchup
eng = api.get_engine(local)
t0 = time.perf_counter()
res = api.simulate_slate(pairs, n=n, seed=5, engine=eng, focus_csv=sheet, usage_dir=gold)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ok = True
if rank == 0:
    sp = priors.load_sp_flex(priors.packaged_priors_path())
    specs = api.slate_specs(pairs, 2 * n, sp, 0, 1)
    for s, (a, b) in zip(specs, pairs):
        s.usage = res[(a, b)]["usage"]
    eng.set_matchups(specs)
    alone = eng.simulate_host(5, want_scores=False, want_hist=True, want_player_hist=True)
    for m, (a, b) in enumerate(pairs):
        e = res[(a, b)]
        ok &= bool(np.array_equal(e["hist"], alone["hist"][m].astype(np.int64)))
        ok &= bool(np.array_equal(e["player_hist"], alone["player_hist"][m].astype(np.int64)))
    top = res[pairs[0]]["props"].head(3).to_dict("records")
    print(json.dumps({"world": world, "games": res["_counters"]["games"], "plays": res["_counters"]["plays"],
                      "merged_equals_single_rank": ok, "seconds": dt, "top_props": top}, default=str), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
