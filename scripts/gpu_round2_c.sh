#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_memo.py -m gpu -x -q -k "not ks_against and not full_size" > gpurun_out/r2_tests_c.log 2>&1; tail -6 gpurun_out/r2_tests_c.log
for t in "24 16" "16 16" "48 16" "24 8" "24 24" "96 28" "12 16"; do
  set -- $t
  FMC_MEMO_TRIPS=$1 FMC_MEMO_BREAK=$2 python scripts/quick_bench.py 4000000 > gpurun_out/r2_sched_$1_$2.log 2>&1; echo "steps $1 break $2"; tail -4 gpurun_out/r2_sched_$1_$2.log
done
python scripts/quick_bench.py 10000000 > gpurun_out/r2_sched_10M.log 2>&1; tail -4 gpurun_out/r2_sched_10M.log
