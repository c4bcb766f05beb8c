"""Quick device-timed throughput probe (no CPU baseline, no e2e): python scripts/quick_bench.py [games] [stage2] [players]
`players`: player mode with the synthetic focus sheet of tests/golden (usage tables, per-game player box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_monte_carlo_b200 import artifacts as art, synth, native
from fast_monte_carlo_b200.engine import Engine, MatchupSpec
games = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
stage2 = sys.argv[2] if len(sys.argv) > 2 else "synthetic"
ms = art.load_default_models()
if stage2 == "synthetic": ms = synth.with_synthetic_stage2(ms)
eng = Engine(ms, stage2="booster" if stage2 == "synthetic" else "standin")
players = len(sys.argv) > 3 and sys.argv[3] == "players"
use = None
if players:
    from fast_monte_carlo_b200 import priors, usage
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    focus = usage.build_focus_usage_tables(os.path.join(gold, "players_focus.csv"))
    sp = priors.load_sp_flex(priors.packaged_priors_path())
    use = tuple(usage.resolve_team(priors.build_team_context_from_sp_flex(t, 2025, 1, sp, focus=focus, usage_dir=gold), ms)
                for t in ("Kansas State", "Iowa State"))
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), games, 0, games, 0, usage=use)])
eng.ctx.packed_slots(0)
box = torch.zeros((games, 2, max(eng.n_slots, 1), 2), dtype=torch.int64, device="cuda") if players else None
cnt = torch.zeros(32, dtype=torch.int64, device="cuda")
hist = torch.zeros((1, 2, 128, 128), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
def step():
    cnt.zero_(); hist.zero_()
    if box is not None: box.zero_()
    eng.ctx.simulate_device(seed=20251018, hist=hist.data_ptr(), counters=cnt.data_ptr(), cuda_stream=st.cuda_stream,
                            players=box.data_ptr() if box is not None else 0)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(3):
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
c = cnt.cpu().numpy()
print(f"rounds/CTA {c[15]/148:.0f} requests {c[16]:.4g} visits/request {c[17]/max(c[16],1):.0f} us/round {best*1e3/(c[15]/148):.1f}")
print(f"memo {eng.memo}: probes {c[20]:.4g} hits {c[21]:.4g} ({c[21]/max(c[20],1):.4f}) walked {c[16]:.4g} trips/warp {c[22]/(148*32):.0f} "
      f"plays/trip/lane {c[1]/max(c[22]*32,1):.3f}")
pr = {"s1": c[3], "s2": c[3] - c[4], "pq": c[4], "rq": c[8], "sq": c[7]}
print("memo hit rate per family: " + "  ".join(f"{k} {c[24 + i] / max(pr[k], 1):.4f}" for i, k in enumerate(pr)) + f"  warp_steps {c[19]:.4g}")
if players: print("player mode: slots per team", eng.n_slots, "packed slots", eng.ctx.packed_slots(0)[:6].tolist())
print(f"{os.environ.get('FMC_LIB_PATH','default')}: {games} games {best:.1f} ms -> {games/best*1e3:.3e} games/s {c[1]/best*1e3:.3e} plays/s  checksum {int(hist.to(torch.int64).mul(torch.arange(128*128*2, device='cuda').view(1,2,128,128)).sum())}")
