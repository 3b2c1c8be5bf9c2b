#!/bin/bash
# 2 GPUs: the sharded bench with the final code (no ride-along at N=2 by the size rule).
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 > gpurun_out/c40_bench_n2.json 2> gpurun_out/c40_bench_n2.err
echo "n2 rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c40_bench_n2.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','sharded_stages_ms_rank0']}, d['e2e']['ms_per_step'], d['factored_root']['ms_per_step'])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/c40_bench_n2.err | tail -3
timeout 300 $TR --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
