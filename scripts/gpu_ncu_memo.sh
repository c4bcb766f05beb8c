#!/bin/bash
# ncu --set full capture of one sim_memo_kernel launch (2 M games of configs[1]) + the remaining GPU tests
set -x
mkdir -p gpurun_out
python scripts/quick_bench.py 2000000 > gpurun_out/r2_quick_pre_ncu.log 2>&1; tail -2 gpurun_out/r2_quick_pre_ncu.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sim_memo -c 1 -o gpurun_out/prof_memo_r2a python scripts/quick_bench.py 2000000 > gpurun_out/r2_ncu_memo.log 2>&1
tail -2 gpurun_out/r2_ncu_memo.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_sim.py > gpurun_out/r2_gputests_rest.log 2>&1; tail -8 gpurun_out/r2_gputests_rest.log
