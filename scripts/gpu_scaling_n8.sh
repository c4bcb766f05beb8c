#!/bin/bash
# 8-GPU validation: headline workload (weak scaling) and the 60-matchup slate (strong scaling)
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/n8_smoke.log 2>&1; tail -1 gpurun_out/n8_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err
cut -c1-300 gpurun_out/n8_bench.json; tail -2 gpurun_out/n8_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload slate --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/n8_slate.json 2> gpurun_out/n8_slate.err
cut -c1-300 gpurun_out/n8_slate.json; tail -2 gpurun_out/n8_slate.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err
cut -c1-300 gpurun_out/n4_bench.json
