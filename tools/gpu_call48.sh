#!/bin/bash
# 1 GPU: full GPU test suite with the line-scatter assemble kernel and the re-mapped panel_cycle kernel, then the step profile.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/c48_tests.txt
timeout 60 python tools/profile_step.py 3 > gpurun_out/c48_profile.log 2>&1; head -16 gpurun_out/step_profile_L3.txt
