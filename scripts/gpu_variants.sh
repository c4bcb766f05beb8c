#!/bin/bash
# usage: gpu_variants.sh name1 name2 ...   (libs in build_variants/libfmc_<name>.so)
mkdir -p gpurun_out
: > gpurun_out/variants.log
for v in "$@"; do
  FMC_LIB_PATH=$PWD/build_variants/libfmc_$v.so python scripts/quick_bench.py 2000000 2>&1 | tail -1 >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
