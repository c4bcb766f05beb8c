#!/bin/bash
# one gpurun call: GPU tests, bench.py (both arms), tree microbench, ncu launch list of the bench command,
# ncu --set full capture of sim_kernel
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -1 gpurun_out/smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1
tail -5 gpurun_out/gputests.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
python bench.py --players --games 4000000 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/bench_players.json 2> gpurun_out/bench_players.err
cat gpurun_out/bench_players.json | cut -c1-400
python scripts/bench_trees.py 67108864 > gpurun_out/trees.json 2> gpurun_out/trees.err
cat gpurun_out/trees.json; tail -3 gpurun_out/trees.err
# launch list of the bench command itself (every kernel once, gpu__time_duration only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/prof_sim_latest python scripts/quick_bench.py 500000 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
