#!/bin/bash
# First GPU call of the next round (run under gpurun, ONE GPU): evidence that round 1 could not afford.
#   gpurun --timeout 1500 -- 'bash tools/ncu_next_round.sh'
# Writes everything to gpurun_out/; summarise into profiles/ afterwards.
set -u
mkdir -p gpurun_out
# 1. tile-configuration sweep of the hoisted GEMM kernel (CUDA events, ~1 min)
timeout 300 ./tools/gemm_lab > gpurun_out/r02_gemm_lab.txt 2>&1
# 2. full ncu capture of the product GEMM kernel on one 8192^3 launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel_hoist -s 2 -c 1 \
    -o gpurun_out/r02_gemm_hoist python tools/one_gemm.py 8192 8192 8192 > gpurun_out/r02_gemm_hoist.log 2>&1
# 3. launch list (durations only) of the adaptive config 5 build, first 600 kernels after the leaf stage
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/r02_config5_launches.csv python tools/run_config5.py --repeat 1 \
    --load-tree tools/data/config5_tree_p10_tol1e-3.npy > gpurun_out/r02_config5_ncu.log 2>&1
# 4. full capture of the adaptive assembly / coarsening kernels (one launch each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"adaptive_gather_kernel|compress_rows_kernel|compress_cols_kernel" \
    -c 6 -o gpurun_out/r02_adaptive_kernels python tools/run_config4.py --p 10 --tol 1e-4 --repeat 1 \
    > gpurun_out/r02_adaptive_kernels.log 2>&1
ls -la gpurun_out
