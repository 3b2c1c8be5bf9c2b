#!/bin/bash
# 1 GPU: phase timing of the block-column kernel (panel_lab), full GPU suite with the structured merges, bench.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 120 ./tools/panel_lab 19200 2 40 > gpurun_out/c20_panel_lab.txt 2>&1
timeout 120 ./tools/panel_lab 19200 2 0 >> gpurun_out/c20_panel_lab.txt 2>&1
timeout 120 ./tools/panel_lab 4800 2 40 >> gpurun_out/c20_panel_lab.txt 2>&1
timeout 120 ./tools/panel_lab 1500 2 40 >> gpurun_out/c20_panel_lab.txt 2>&1
cat gpurun_out/c20_panel_lab.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/c20_pytest.log
timeout 600 python bench.py --steps 2 --warmup 2 --factored 0 --host-resident 0 > gpurun_out/c20_bench_n1.json 2> gpurun_out/c20_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c20_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, {k:d['stages'][k] for k in ['local_solve_ms','merge_ms','down_pass_ms']}, d['e2e']['ms_per_step'])
r=d['roofline']; print({k:r.get(k) for k in ['achieved','frac','gemm_ms_per_step','other_kernels_ms_per_step']})
PY
