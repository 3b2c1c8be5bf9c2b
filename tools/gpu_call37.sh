#!/bin/bash
# 1 GPU: narrow right-hand sides on the look-ahead stream: full suite, bench.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c37_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c37_pytest.log
timeout 600 python bench.py --host-resident 0 --factored 0 > gpurun_out/c37_bench_n1.json 2> gpurun_out/c37_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c37_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, {k:d['stages'][k] for k in ['local_solve_ms','merge_ms','down_pass_ms']}, d['e2e']['ms_per_step'])
PY
