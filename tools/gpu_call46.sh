#!/bin/bash
# 1 GPU: ncu of the paired transposed mat-vec kernel at the root size.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"gemv_t_pair_kernel" -s 2 -c 1 -o gpurun_out/c46_gemv_t_pair python tools/one_gemv_t.py > gpurun_out/c46_ncu.log 2>&1; echo "ncu rc=$?"
