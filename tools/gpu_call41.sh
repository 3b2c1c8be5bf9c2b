#!/bin/bash
# 1 GPU: transposed mat-vec primitive test, host-resident e2e through pinned result buffers.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_e2e.py -m gpu -q > gpurun_out/c41_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c41_pytest.log
timeout 600 python bench.py --factored 0 > gpurun_out/c41_bench_n1.json 2> gpurun_out/c41_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c41_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, d['e2e']['ms_per_step'], d['e2e_host_resident'])
PY
tail -3 gpurun_out/c41_bench_n1.err
