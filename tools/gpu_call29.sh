#!/bin/bash
# 1 GPU: adaptive down pass replayed as a CUDA graph: adaptive GPU tests, config 5 solve timings.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_adaptive.py tests/test_gpu_baseline_configs.py -m gpu -q > gpurun_out/c29_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c29_pytest.log
timeout 600 python tools/run_config5.py --repeat 1 --load-tree tools/data/config5_tree_p10_tol1e-3.npy > gpurun_out/c29_config5.log 2>&1; tail -1 gpurun_out/c29_config5.log | cut -c1-1200
