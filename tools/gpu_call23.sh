#!/bin/bash
# 1 GPU: register-panel diagonal-block kernel + paired triangular inversions: GPU suite, LU micro-benchmark, bench.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c23_pytest.log
for cfg in "19200 1 0" "4800 8 9601"; do
  for db in reg smem; do
    echo "== $cfg diagblk=$db"; HPS_LU_FORCE_SPEC=1 HPS_DIAGBLK=$db timeout 300 python tools/bench_lu.py $cfg 2 2>&1 | tail -4
  done
done > gpurun_out/c23_bench_lu.txt 2>&1
grep -E "==|iter 2|per category|residual|Error|error" gpurun_out/c23_bench_lu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --host-resident 0 --factored 0 > gpurun_out/c23_bench_n1.json 2> gpurun_out/c23_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c23_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, {k:d['stages'][k] for k in ['local_solve_ms','merge_ms','down_pass_ms']}, d['e2e']['ms_per_step'])
r=d['roofline']; print({k:r.get(k) for k in ['achieved','frac','gemm_ms_per_step','other_kernels_ms_per_step']})
PY
