#!/bin/bash
# one gpurun call for the player mode: every GPU test, production and player-mode throughput, bench.py --players,
# ncu launch list + one --set full capture of the player instantiation of sim_kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/p_tests.log 2>&1
tail -8 gpurun_out/p_tests.log
: > gpurun_out/p_quick.log
python scripts/quick_bench.py 2000000 2>&1 | tail -1 >> gpurun_out/p_quick.log
python scripts/quick_bench.py 2000000 synthetic players 2>&1 | tail -3 >> gpurun_out/p_quick.log
cat gpurun_out/p_quick.log
python bench.py --players --games 4000000 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
cat gpurun_out/p_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/p_launches.csv python bench.py --players --games 1000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/p_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/prof_sim_players python scripts/quick_bench.py 500000 synthetic players > gpurun_out/p_ncu.log 2>&1
tail -2 gpurun_out/p_ncu.log
