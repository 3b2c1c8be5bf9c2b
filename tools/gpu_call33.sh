#!/bin/bash
# 4 GPUs: sharded bench at N=4.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/c33_bench_n4.json 2> gpurun_out/c33_bench_n4.err
echo "n4 rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c33_bench_n4.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','sharded_stages_ms_rank0']}, d['e2e']['ms_per_step'], d['factored_root']['ms_per_step'])
print(d['roofline']['other_kernels_ms_per_step'])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/c33_bench_n4.err | tail -3
