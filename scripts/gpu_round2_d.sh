#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 60 python scripts/dbg_variant.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_sim.py tests/test_gpu_memo.py -m gpu -x -q -k "not ks_against and not full_size" > gpurun_out/r2_tests_d.log 2>&1; tail -6 gpurun_out/r2_tests_d.log
for t in "6 8" "1 1" "4 6" "8 12" "12 16" "6 1" "32 32"; do
  set -- $t
  FMC_MEMO_MIN_RARE=$1 FMC_MEMO_MIN_S2=$2 timeout 120 python scripts/quick_bench.py 4000000 > gpurun_out/r2_rare_$1_$2.log 2>&1; echo "min_rare $1 min_s2 $2"; tail -4 gpurun_out/r2_rare_$1_$2.log
done
FMC_MEMO_TRIPS=16 timeout 120 python scripts/quick_bench.py 4000000 2>&1 | tail -1
FMC_MEMO_TRIPS=5 timeout 120 python scripts/quick_bench.py 4000000 2>&1 | tail -1
timeout 120 python scripts/quick_bench.py 10000000 > gpurun_out/r2_rare_10M.log 2>&1; tail -4 gpurun_out/r2_rare_10M.log
