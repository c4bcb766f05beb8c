#!/bin/bash
set -x
mkdir -p gpurun_out
bash scripts/gpu_variants.sh 4000000 base nosw t512x2 t256x4 t768x1 env:FMC_MEMO_TRIPS=16:t512x2 env:FMC_MEMO_TRIPS=16:t256x4
cp gpurun_out/variants.log gpurun_out/r2_variants_b.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sim_memo -c 1 -o gpurun_out/prof_memo_r2b python scripts/quick_bench.py 4000000 > gpurun_out/r2_ncu_memo_b.log 2>&1
tail -3 gpurun_out/r2_ncu_memo_b.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_all.log 2>&1; tail -8 gpurun_out/r2_gputests_all.log
