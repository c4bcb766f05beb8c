#!/bin/bash
# first GPU contact of the memo kernel: smoke, the sim tests, quick throughput probes with the memo off / on
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests/test_gpu_sim.py -m gpu -x -q > gpurun_out/r2_simtests.log 2>&1; tail -8 gpurun_out/r2_simtests.log
FMC_MEMO=off python scripts/quick_bench.py 2000000 > gpurun_out/r2_quick_off.log 2>&1; tail -3 gpurun_out/r2_quick_off.log
for t in "8 16" "4 16" "16 16" "8 8" "8 24" "32 28"; do
  set -- $t
  FMC_MEMO=on FMC_MEMO_TRIPS=$1 FMC_MEMO_BREAK=$2 python scripts/quick_bench.py 2000000 > gpurun_out/r2_quick_on_$1_$2.log 2>&1; echo "trips $1 break $2"; tail -3 gpurun_out/r2_quick_on_$1_$2.log
done
FMC_MEMO=on python scripts/quick_bench.py 10000000 > gpurun_out/r2_quick_on_10M.log 2>&1; tail -3 gpurun_out/r2_quick_on_10M.log
