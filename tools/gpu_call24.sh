#!/bin/bash
# 1 GPU: no-search diagonal-block kernel + register-resident triangular inversions: GPU suite, LU micro-benchmarks.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c24_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c24_pytest.log
for cfg in "19200 1 0" "4800 8 9601" "1000 512 601"; do
  for db in nosearch pivot; do
    echo "== $cfg diagblk=$db"; HPS_LU_FORCE_SPEC=1 HPS_DIAGBLK=$db timeout 300 python tools/bench_lu.py $cfg 2 2>&1 | tail -4
  done
done > gpurun_out/c24_bench_lu.txt 2>&1
grep -E "==|iter 2|per category|residual|Error|error" gpurun_out/c24_bench_lu.txt
timeout 300 python tools/lu_accuracy.py > gpurun_out/c24_lu_accuracy.txt 2>&1; tail -12 gpurun_out/c24_lu_accuracy.txt
