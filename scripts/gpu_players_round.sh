#!/bin/bash
# one gpurun call for the player mode: its GPU tests, the simulation regression tests, production throughput
# of the previous vs the current library, player-mode throughput
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_players.py -x -q > gpurun_out/p_tests.log 2>&1
tail -15 gpurun_out/p_tests.log
python -m pytest tests/test_gpu_sim.py -x -q > gpurun_out/p_simtests.log 2>&1
tail -3 gpurun_out/p_simtests.log
: > gpurun_out/p_quick.log
[ -f build_variants/libfmc_prev.so ] && FMC_LIB_PATH=$PWD/build_variants/libfmc_prev.so python scripts/quick_bench.py 2000000 2>&1 | tail -1 >> gpurun_out/p_quick.log
python scripts/quick_bench.py 2000000 2>&1 | tail -1 >> gpurun_out/p_quick.log
python scripts/quick_bench.py 1000000 synthetic players 2>&1 | tail -3 >> gpurun_out/p_quick.log
cat gpurun_out/p_quick.log
