#!/bin/bash
# 8 GPUs: sharded bench at N=8 (L=3 + the L=4 target) and N=4.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/c5_bench_n8.json 2> gpurun_out/c5_bench_n8.err
echo "n8 rc=$?"; cat gpurun_out/c5_bench_n8.json; grep -v "^\*\|OMP_NUM" gpurun_out/c5_bench_n8.err | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/c5_bench_n4.json 2> gpurun_out/c5_bench_n4.err
echo "n4 rc=$?"; cat gpurun_out/c5_bench_n4.json; grep -v "^\*\|OMP_NUM" gpurun_out/c5_bench_n4.err | tail -5
