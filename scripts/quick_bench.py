"""Quick device-timed throughput probe (no CPU baseline, no e2e): python scripts/quick_bench.py [games] [stage2]"""
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
eng.set_matchups([MatchupSpec("Kansas State", "Iowa State", (15.6, 35.7, 20.0), (11.0, 31.5, 20.6), games, 0, games, 0)])
eng.ctx.packed_slots(0)
cnt = torch.zeros(32, dtype=torch.int64, device="cuda")
hist = torch.zeros((1, 2, 128, 128), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
def step():
    cnt.zero_(); hist.zero_()
    eng.ctx.simulate_device(seed=20251018, hist=hist.data_ptr(), counters=cnt.data_ptr(), cuda_stream=st.cuda_stream)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(3):
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
c = cnt.cpu().numpy()
print(f"rounds/CTA {c[15]/148:.0f} requests {c[16]:.4g} visits/request {c[17]/max(c[16],1):.0f} us/round {best*1e3/(c[15]/148):.1f}")
print(f"{os.environ.get('FMC_LIB_PATH','default')}: {games} games {best:.1f} ms -> {games/best*1e3:.3e} games/s {c[1]/best*1e3:.3e} plays/s  checksum {int(hist.to(torch.int64).mul(torch.arange(128*128*2, device='cuda').view(1,2,128,128)).sum())}")
