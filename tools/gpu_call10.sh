#!/bin/bash
# 1 GPU: after row equilibration of the ItI leaf systems: ItI accuracy study, full GPU suite.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/iti_accuracy.py > gpurun_out/c10_iti_accuracy.txt 2>&1; tail -12 gpurun_out/c10_iti_accuracy.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
grep -E "config 2|arbitration|passed|failed|FAILED|rc=|Error" gpurun_out/c10_pytest.log | tail -25
