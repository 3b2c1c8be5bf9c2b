#!/bin/bash
# 1 GPU: full GPU suite after the blocked-substitution upper inverse, step profile (stage windows, GEMM shapes).
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
grep -E "config 2|arbitration|passed|failed|FAILED|rc=|Error" gpurun_out/c16_pytest.log | tail -12
timeout 500 python tools/profile_step.py 3 > gpurun_out/c16_profile.log 2>&1; grep -E "^L=|^GEMM|K<|K>|^window|trtri" gpurun_out/c16_profile.log | head -40
