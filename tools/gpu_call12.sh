#!/bin/bash
# 4 GPUs: distributed-LU micro-benchmark with per-launch timelines of every rank; then the sharded bench at N=4.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tools/bench_dist_lu.py 19200 9600 2 > gpurun_out/c12_dist_lu.txt 2>&1
grep -E "world=|rank 0|Error|error" gpurun_out/c12_dist_lu.txt
timeout 600 $TR --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/c12_bench_n4.json 2> gpurun_out/c12_bench_n4.err
echo "n4 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c12_bench_n4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, d['roofline']['other_kernels_ms_per_step'])
PY
