#!/bin/bash
# 1 GPU: final validation of the round: full GPU suite, smoke, default bench line, step profile.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c42_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c42_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/c42_bench_n1.json 2> gpurun_out/c42_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c42_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','gpu_launches']}, d['e2e']['ms_per_step'], d['e2e_host_resident']['ms_per_step'], d['factored_root']['ms_per_step'])
print(d['roofline']['hbm_kernels'], d['roofline']['frac'], d['stages']['local_solve_frac_of_fp64_peak'], d['stages']['down_pass_frac_of_hbm_peak'])
PY
timeout 300 python tools/profile_step.py 3 > gpurun_out/c42_profile.log 2>&1; grep -E "^L=|^GEMM|K<|K>|stream 0: busy" gpurun_out/c42_profile.log | head
