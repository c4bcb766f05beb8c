#!/bin/bash
# round-2 validation on one GPU: full GPU test-suite, bench (both arms), predict ncu, slate + season (configs[3], [4]) at N=1,
# the G6 record, the coherent-lane walk microbench
set -x
K="timeout -s KILL"
mkdir -p gpurun_out
$K 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_all.log 2>&1; tail -5 gpurun_out/r2_gputests_all.log
$K 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_bench.json; tail -2 gpurun_out/r2_bench.err
$K 120 ./build_variants/walk_mb > gpurun_out/r2_walk_microbench.log 2>&1; head -34 gpurun_out/r2_walk_microbench.log
$K 400 ncu --set full --clock-control none -k regex:predict_kernel -c 4 -o gpurun_out/prof_predict_r2c python scripts/bench_trees.py 4194304 > gpurun_out/r2_ncu_predict_c.log 2>&1; tail -2 gpurun_out/r2_ncu_predict_c.log | cut -c1-300
$K 400 python bench.py --workload slate --slate-games 1000000 --steps 2 --warmup 3 --e2e-steps 1 --no-roofline --no-tree-eval --no-cpu-baseline > gpurun_out/r2_slate60x1M_n1.json 2> gpurun_out/r2_slate_n1.err; cut -c1-250 gpurun_out/r2_slate60x1M_n1.json; tail -2 gpurun_out/r2_slate_n1.err
$K 400 python bench.py --workload season --slate-games 100000 --steps 2 --warmup 3 --e2e-steps 1 --no-roofline --no-tree-eval --no-cpu-baseline > gpurun_out/r2_season720x100k_n1.json 2> gpurun_out/r2_season_n1.err; cut -c1-250 gpurun_out/r2_season720x100k_n1.json; tail -2 gpurun_out/r2_season_n1.err
$K 600 python scripts/g6_full.py 1000000 gpurun_out/r2_g6_1M.json 2>&1 | tail -2
