#!/bin/bash
# one gpurun call for the player mode: its GPU tests (+ API tests), production and player-mode throughput
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_players.py tests/test_gpu_api.py -x -q > gpurun_out/p_tests.log 2>&1
tail -15 gpurun_out/p_tests.log
: > gpurun_out/p_quick.log
python scripts/quick_bench.py 2000000 2>&1 | tail -1 >> gpurun_out/p_quick.log
python scripts/quick_bench.py 2000000 synthetic players 2>&1 | tail -3 >> gpurun_out/p_quick.log
cat gpurun_out/p_quick.log
