#!/bin/bash
# 2 GPUs: distributed root with speculative block columns, structured ride-along right-hand sides and balanced panels.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tools/bench_dist_lu.py 19200 19200 2 > gpurun_out/c22_dist_lu.txt 2>&1
grep -E "world=|rank 0|Error|error" gpurun_out/c22_dist_lu.txt
timeout 900 $TR --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c22_bench_n2.json 2> gpurun_out/c22_bench_n2.err
echo "n2 rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c22_bench_n2.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, d['e2e']['ms_per_step'], d['factored_root'], d['roofline']['other_kernels_ms_per_step'])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/c22_bench_n2.err | tail -5
