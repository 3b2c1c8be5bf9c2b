#!/bin/bash
# 1 GPU: forward substitution after the distributed factorisation (no ride-along) vs ride-along, full suite, smoke.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c34_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c34_pytest.log
HPS_DIST_RIDE=1 timeout 600 python -m pytest tests/test_gpu_config3.py tests/test_gpu_stages.py -m gpu -q -k "sharded or distributed or factored" > gpurun_out/c34_pytest_ride.log 2>&1; echo "ride pytest rc=$?"; tail -3 gpurun_out/c34_pytest_ride.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for ride in 0 1; do echo "== single-rank distributed root, ride=$ride"; HPS_DIST_RIDE=$ride timeout 300 python tools/bench_dist_lu.py 19200 9600 1 2>&1 | grep -E "iter 1|rank 0 per" | cut -c1-400; done
