#!/bin/bash
# 1 GPU: full GPU suite on the restored tree, the default bench line, the reference arm, step profile.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
tail -5 gpurun_out/c17_pytest.log
timeout 600 python bench.py > gpurun_out/c17_bench_n1.json 2> gpurun_out/c17_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c17_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c17_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','stages','e2e','same_config_sample']})
r=d['roofline']; print({k:r.get(k) for k in ['achieved','peak','frac','gemm_ms_per_step','other_kernels_ms_per_step']})
PY
timeout 500 python tools/profile_step.py 3 > gpurun_out/c17_profile.log 2>&1; grep -E "^L=|^GEMM|K<|K>|^window|trtri|panel" gpurun_out/c17_profile.log | head -60
