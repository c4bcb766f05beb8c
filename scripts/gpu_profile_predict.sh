#!/bin/bash
# ncu of the tree-predict kernels (tree microbench, BASELINE configs[2]) + launch list
set -x
mkdir -p gpurun_out
python scripts/bench_trees.py 4194304 4194304 > gpurun_out/d_trees_small.json 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -c 4 -o gpurun_out/prof_predict_r01 python scripts/bench_trees.py 4194304 4194304 > gpurun_out/d_ncu_predict.log 2>&1
tail -2 gpurun_out/d_ncu_predict.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/d_launches_trees.csv python scripts/bench_trees.py 4194304 4194304 > /dev/null 2>&1
