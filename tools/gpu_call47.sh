#!/bin/bash
# 1 GPU: ncu --set full of assemble_kernel<3> and the three panel_cycle_kernel launches of one L=3 build.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"assemble_kernel|panel_cycle_kernel" -c 4 -o gpurun_out/c47_hbm python tools/time_stages.py 3 > gpurun_out/c47_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/c47_ncu.log
