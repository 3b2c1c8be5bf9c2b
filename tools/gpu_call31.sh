#!/bin/bash
# 1 GPU: threshold-pivoting speculation in the DtN leaf solves: full suite, bench, accuracy of the leaf stage.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c31_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/c31_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --host-resident 0 --factored 0 > gpurun_out/c31_bench_n1.json 2> gpurun_out/c31_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c31_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, {k:d['stages'][k] for k in ['local_solve_ms','merge_ms','down_pass_ms','local_solve_frac_of_fp64_peak']}, d['e2e']['ms_per_step'])
r=d['roofline']; print({k:r.get(k) for k in ['achieved','frac','gemm_ms_per_step','other_kernels_ms_per_step']})
PY
timeout 300 python tools/profile_step.py 3 > gpurun_out/c31_profile.log 2>&1; grep -E "^window 0" gpurun_out/c31_profile.log | cut -c1-600
