#!/bin/bash
# 1 GPU: bandwidth of the transposed mat-vec at the root size; ncu of panel_cycle_kernel (root launch) and gemv_t_kernel.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 120 python tools/one_gemv_t.py 2>&1 | tail -5
timeout 300 ncu --set full --clock-control none -k regex:"gemv_t_kernel" -s 2 -c 1 -o gpurun_out/c44_gemv_t python tools/one_gemv_t.py > gpurun_out/c44_ncu_gemv_t.log 2>&1; echo "ncu1 rc=$?"
ls -la gpurun_out/*.ncu-rep
