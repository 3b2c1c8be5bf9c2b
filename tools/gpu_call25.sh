#!/bin/bash
# 8 GPUs: distributed-LU micro-benchmark and the sharded bench at N=8 (L=3 + the L=4 target) with the speculative
# block columns, structured right-hand sides and balanced panels.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tools/bench_dist_lu.py 19200 4800 2 > gpurun_out/c25_dist_lu.txt 2>&1
grep -E "world=|rank 0|Error|error" gpurun_out/c25_dist_lu.txt
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/c25_bench_n8.json 2> gpurun_out/c25_bench_n8.err
echo "n8 rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c25_bench_n8.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','sharded_stages_ms_rank0']}, d['e2e']['ms_per_step'], d['factored_root'])
print(d['roofline']['other_kernels_ms_per_step'])
print(d['target_L4'])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/c25_bench_n8.err | tail -5
