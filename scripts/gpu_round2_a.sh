#!/bin/bash
set -x
mkdir -p gpurun_out
for v in t512lb t512o1 t992; do echo == $v; CUDA_LAUNCH_BLOCKING=1 FMC_LIB_PATH=$PWD/build_variants/libfmc_$v.so python scripts/dbg_variant.py 2>&1 | tail -3; done > gpurun_out/r2_bisect.log 2>&1
cat gpurun_out/r2_bisect.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_all.log 2>&1; tail -8 gpurun_out/r2_gputests_all.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cat gpurun_out/r2_bench.json | cut -c1-1500; tail -3 gpurun_out/r2_bench.err
