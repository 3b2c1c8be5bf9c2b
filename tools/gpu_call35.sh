#!/bin/bash
# 8 GPUs: the L=4 target in isolation, forward substitution after the factorisation (auto) vs ride-along.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29541 tools/run_L4.py 1 > gpurun_out/c35_L4_auto.log 2>&1; echo "auto rc=$?"; grep "^{" gpurun_out/c35_L4_auto.log
HPS_DIST_RIDE=1 timeout 400 $TR --master-port 29542 tools/run_L4.py 1 > gpurun_out/c35_L4_ride.log 2>&1; echo "ride rc=$?"; grep "^{" gpurun_out/c35_L4_ride.log
grep -E "Error|error" gpurun_out/c35_L4_*.log | head -5
