#!/bin/bash
# 2 GPUs: GPU suite (new gather / assemble kernels), then the sharded bench with the P2P root factorisation and
# with the NCCL-broadcast fallback.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -4 gpurun_out/c4_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/c4_bench_n2_p2p.json 2> gpurun_out/c4_bench_n2_p2p.err
echo "p2p rc=$?"; cat gpurun_out/c4_bench_n2_p2p.json; tail -5 gpurun_out/c4_bench_n2_p2p.err
HPS_DIST_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/c4_bench_n2_nccl.json 2> gpurun_out/c4_bench_n2_nccl.err
echo "nccl rc=$?"; cat gpurun_out/c4_bench_n2_nccl.json; tail -3 gpurun_out/c4_bench_n2_nccl.err
