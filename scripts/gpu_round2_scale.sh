#!/bin/bash
# usage: gpu_round2_scale.sh N what...   (what: headline slate season)   -- run under gpurun --gpus N
set -x
N=$1; shift
K="timeout -s KILL"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for w in "$@"; do
  case $w in
    headline) $K 420 $TR bench.py --gpus $N --no-tree-eval > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; cut -c1-260 gpurun_out/r2_bench_n$N.json; tail -2 gpurun_out/r2_bench_n$N.err;;
    slate) $K 300 $TR bench.py --gpus $N --workload slate --slate-games 1000000 --steps 2 --warmup 3 --e2e-steps 1 --no-roofline --no-tree-eval --no-cpu-baseline > gpurun_out/r2_slate60x1M_n$N.json 2> gpurun_out/r2_slate_n$N.err; cut -c1-260 gpurun_out/r2_slate60x1M_n$N.json; tail -2 gpurun_out/r2_slate_n$N.err;;
    season) $K 300 $TR bench.py --gpus $N --workload season --slate-games 100000 --steps 2 --warmup 3 --e2e-steps 1 --no-roofline --no-tree-eval --no-cpu-baseline > gpurun_out/r2_season720x100k_n$N.json 2> gpurun_out/r2_season_n$N.err; cut -c1-260 gpurun_out/r2_season720x100k_n$N.json; tail -2 gpurun_out/r2_season_n$N.err;;
  esac
done
