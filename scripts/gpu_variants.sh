#!/bin/bash
# usage: gpu_variants.sh games name1 name2 ...   (libs in build_variants/libfmc_<name>.so; name "env:K=V:name" sets an env var)
games=$1; shift
mkdir -p gpurun_out
: > gpurun_out/variants.log
for v in "$@"; do
  envs=""
  while [[ "$v" == env:* ]]; do v=${v#env:}; envs="$envs ${v%%:*}"; v=${v#*:}; done
  echo "== $v $envs" >> gpurun_out/variants.log
  env $envs FMC_LIB_PATH=$PWD/build_variants/libfmc_$v.so python scripts/quick_bench.py $games 2>&1 | tail -2 >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
