#!/bin/bash
# 1 GPU: transposed mat-vec after the unroll / wider grid: bandwidth, then the tests that use it.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 120 python tools/one_gemv_t.py 2>&1 | tail -5 | tee gpurun_out/c45_gemv_t.txt
timeout 240 python -m pytest tests/test_gpu_adjoint.py tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/c45_tests.txt
