#!/bin/bash
# 1 GPU: validation after pruning the superseded kernel variants: full suite, smoke, default bench.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c36_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c36_pytest.log
HPS_LU_SPEC=0 timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_config3.py -m gpu -q > gpurun_out/c36_pytest_pivoted.log 2>&1; echo "pivoted-path pytest rc=$?"; tail -3 gpurun_out/c36_pytest_pivoted.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/c36_bench_n1.json 2> gpurun_out/c36_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c36_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','executed_gemm_tflop_per_step']}, {k:d['stages'][k] for k in ['local_solve_ms','merge_ms','down_pass_ms']}, d['e2e']['ms_per_step'], d['e2e_host_resident']['ms_per_step'], d['factored_root']['ms_per_step'], d['same_config_sample']['e2e_leaves_per_s'], d['cpu_baseline']['value'])
print(d['clocks'], d['gpu_launches'])
PY
