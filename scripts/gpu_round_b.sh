#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/b_gputests.log 2>&1
tail -5 gpurun_out/b_gputests.log
python scripts/quick_bench.py 2000000 > gpurun_out/b_quick.log 2>&1
for v in nopf; do
  FMC_LIB_PATH=$PWD/build_variants/libfmc_$v.so python scripts/quick_bench.py 2000000 >> gpurun_out/b_quick.log 2>&1
done
python scripts/quick_bench.py 2000000 standin >> gpurun_out/b_quick.log 2>&1
cat gpurun_out/b_quick.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/prof_sim_r01d python scripts/quick_bench.py 500000 > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
