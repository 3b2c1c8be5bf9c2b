#!/bin/bash
# 1 GPU: default bench line with the other BASELINE configs embedded.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/c43_bench_n1.json 2> gpurun_out/c43_bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"
python - <<PY
import json
d=json.loads(open('gpurun_out/c43_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, d['e2e']['ms_per_step'])
print(json.dumps(d['other_configs'], indent=1))
PY
tail -3 gpurun_out/c43_bench_n1.err
