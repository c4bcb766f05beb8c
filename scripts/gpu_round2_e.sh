#!/bin/bash
# validation of the final kernels: tests under both memo settings, bench (both arms), ncu launch list + full captures
set -x
K="timeout -s KILL"
mkdir -p gpurun_out
$K 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
$K 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_all.log 2>&1; tail -6 gpurun_out/r2_gputests_all.log
FMC_MEMO=off $K 600 python -m pytest tests/test_gpu_sim.py tests/test_gpu_players.py -m gpu -x -q -k "not ks_against and not full_size" > gpurun_out/r2_gputests_memo_off.log 2>&1; tail -3 gpurun_out/r2_gputests_memo_off.log
$K 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-400 gpurun_out/r2_bench.json; tail -2 gpurun_out/r2_bench.err
$K 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-300 gpurun_out/r2_bench_ref.json
$K 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tree-states 4194304 --roofline-steps 1 > gpurun_out/r2_ncu_launch.log 2>&1; tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-300
$K 600 ncu --set full --clock-control none --import-source on -k regex:sim_memo -c 1 -o gpurun_out/prof_memo_r2c python scripts/quick_bench.py 4000000 > gpurun_out/r2_ncu_memo_c.log 2>&1; tail -4 gpurun_out/r2_ncu_memo_c.log
FMC_MEMO=off $K 600 ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/prof_sim_r2c python scripts/quick_bench.py 500000 > gpurun_out/r2_ncu_sim_c.log 2>&1; tail -4 gpurun_out/r2_ncu_sim_c.log
$K 600 ncu --set full --clock-control none -k regex:predict_kernel -c 4 -o gpurun_out/prof_predict_r2c python scripts/bench_trees.py 4194304 > gpurun_out/r2_ncu_predict_c.log 2>&1; tail -2 gpurun_out/r2_ncu_predict_c.log | cut -c1-300
