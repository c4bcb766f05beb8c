"""Tree-inference microbench (BASELINE.json configs[2]): play_model.xgb + the pass / run /
sack q10/q50/q90 models (4,600 trees) on synthetic game states, through fmc_tree_predict (device buffers).

    python scripts/bench_trees.py [n_states=2**26] [batch=2**22]

States follow SURVEY 8(d): down p=(.38,.31,.21,.10), distance=clip(round(N(8,4),1),.5,30), yardsToGoal U{1..99},
score_diff round(N(0,14)), seconds U{1..3600}, timeouts 3/3, (off, def) a random ordered pair of the 136 CSV
teams, flags derived, names "Unknown", no coach.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fast_monte_carlo_b200 import artifacts as art, native, priors
from fast_monte_carlo_b200.engine import Engine

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
ms = art.load_default_models()
eng = Engine(ms, stage2="standin")
sp = priors.load_sp_flex(priors.packaged_priors_path()).drop_duplicates(subset=["RATING", "OFFENSE", "DEFENSE"])
teams = sp[["RATING", "OFFENSE", "DEFENSE"]].to_numpy(dtype=float)


def make_states(n, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 17))
    x[:, 0] = rng.choice([1, 2, 3, 4], size=n, p=[.38, .31, .21, .10])
    x[:, 1] = np.clip(np.round(rng.normal(8, 4, n), 1), 0.5, 30)
    x[:, 2] = rng.integers(1, 100, n)
    x[:, 3] = x[:, 2] <= 20
    x[:, 4] = np.round(rng.normal(0, 14, n))
    x[:, 5] = rng.integers(1, 3601, n)
    x[:, 6] = x[:, 7] = 3
    if teams is not None:
        o = rng.integers(0, len(teams), n); d = (o + rng.integers(1, len(teams), n)) % len(teams)
        x[:, 8] = teams[o, 0]; x[:, 9] = teams[o, 1]; x[:, 10] = teams[d, 2]; x[:, 11] = teams[d, 0]
    else:
        x[:, 8:12] = np.round(rng.normal(5, 12, (n, 4)), 1)
    x[:, 12] = x[:, 1] >= x[:, 2] - 0.5
    x[:, 13] = (x[:, 0] == 4) & (x[:, 1] <= 2)
    x[:, 14] = x[:, 2] <= 33
    x[:, 15] = np.where(x[:, 5] > 1800, 1, 2)
    x[:, 16] = (x[:, 5] % 1800) <= 120
    return x


rows = torch.from_numpy(make_states(batch, 3)).cuda()
models = [(art.MODEL_IDS[name], name) for name in ("play_model", "pass_yards", "run_yards", "sack_yards")]
outs = {mid: torch.empty((batch, ms[name].n_outputs), dtype=torch.float64, device="cuda") for mid, name in models}
st = torch.cuda.current_stream()


def run_batch():
    for mid, name in models:
        eng.ctx.tree_predict_device(mid, rows.data_ptr(), batch, outs[mid].data_ptr(), cuda_stream=st.cuda_stream)


run_batch(); torch.cuda.synchronize()
n_batches = max(1, n_total // batch)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n_batches):
    run_batch()
e1.record(); torch.cuda.synchronize()
msec = e0.elapsed_time(e1)
n_done = n_batches * batch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
algo = sum(bench.algorithmic_bytes_per_row(ms[name]) for _, name in models)
print(json.dumps({
    "metric": "tree_microbench_states_per_sec", "value": n_done / msec * 1e3, "unit": "states/s", "states": n_done,
    "ms": msec, "trees_per_state": sum(ms[name].n_trees for _, name in models),
    "algorithmic_bytes_per_state": algo, "algorithmic_gbs": algo * n_done / msec * 1e3 / 1e9,
    "models": [name for _, name in models], "batch": batch,
    "note": "fmc_tree_predict on resident float64 [n][17] rows, general rows (nothing folded but the one-hots); "
            "packed tables are cached per model in the context (packed + uploaded once, before the timed region)",
    "checksum": float(sum(o.sum().item() for o in outs.values())),
}))
