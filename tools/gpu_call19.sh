#!/bin/bash
# 1 GPU: fence-free block-column kernel (A/B against the fenced one), structured forward substitution of the merges
# (A/B with HPS_MERGE_STRUCT=0), full GPU suite.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c19_pytest.log
for cfg in "19200 1 0" "4800 1 0" "1200 8 0" "2400 4 0"; do
  for mode in tagged fenced; do
    echo "== $cfg blockcol=$mode"; HPS_BLOCKCOL=$mode timeout 300 python tools/bench_lu.py $cfg 2 2>&1 | tail -3
  done
done > gpurun_out/c19_bench_lu.txt 2>&1
grep -E "==|iter 2|per category" gpurun_out/c19_bench_lu.txt
for st in 1 0; do
  HPS_MERGE_STRUCT=$st timeout 600 python bench.py --steps 2 --warmup 2 --factored 0 --host-resident 0 > gpurun_out/c19_bench_n1_struct$st.json 2> gpurun_out/c19_bench_n1_struct$st.err; echo "bench struct=$st rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/c19_bench_n1_struct$st.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, {k:d['stages'][k] for k in ['local_solve_ms','merge_ms','down_pass_ms']}, d['e2e']['ms_per_step'])
r=d['roofline']; print({k:r.get(k) for k in ['achieved','frac','gemm_ms_per_step','other_kernels_ms_per_step']})
PY
done
timeout 300 python tools/profile_step.py 3 > gpurun_out/c19_profile.log 2>&1; grep -E "^L=|^GEMM|K<|K>|^window" gpurun_out/c19_profile.log | head -20
