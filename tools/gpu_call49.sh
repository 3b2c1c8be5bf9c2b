#!/bin/bash
# 1 GPU: short bench line (no factored / host-resident / other-config extras) after the assemble and panel-permutation kernels.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 50 python bench.py --steps 3 --warmup 3 --factored 0 --host-resident 0 --other-configs 0 > gpurun_out/c49_bench.json 2> gpurun_out/c49_bench.err; echo "rc=$?"
tail -c 600 gpurun_out/c49_bench.json
