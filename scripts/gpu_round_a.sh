#!/bin/bash
# one GPU call: parity tests, shape/ILP variants, ncu capture of the current kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_gputests.log 2>&1
tail -3 gpurun_out/a_gputests.log
python scripts/quick_bench.py 2000000 > gpurun_out/a_quick.log 2>&1
for v in t256x4 t512x2 t768x1 ilp2 ilp6 ilp8; do
  FMC_LIB_PATH=$PWD/build_variants/libfmc_$v.so python scripts/quick_bench.py 2000000 >> gpurun_out/a_quick.log 2>&1
done
python scripts/quick_bench.py 2000000 standin >> gpurun_out/a_quick.log 2>&1
cat gpurun_out/a_quick.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/prof_sim_r01c python scripts/quick_bench.py 100000 > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_ncu.log
