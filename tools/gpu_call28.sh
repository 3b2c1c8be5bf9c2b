#!/bin/bash
# 1 GPU: adjoint (incl. top-operator tangent), device mesh refinement check, full suite.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/c28_pytest.log
for md in none cuda:0; do
  if [ "$md" = none ]; then a=""; else a="--mesh-device $md"; fi
  timeout 400 python tools/run_config4.py --p 10 --tol 1e-5 --repeat 1 $a > gpurun_out/c28_config4_$md.log 2>&1; echo "mesh device $md:"; tail -1 gpurun_out/c28_config4_$md.log | cut -c1-400
done
