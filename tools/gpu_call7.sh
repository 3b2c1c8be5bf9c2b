#!/bin/bash
# 1 GPU: full GPU suite (scattering, interpolation, factored root, config 3), ItI accuracy study, bench with factored mode.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
grep -E "config 3|passed|failed|FAILED|rc=|Error" gpurun_out/c7_pytest.log | tail -25
timeout 600 python tools/iti_accuracy.py > gpurun_out/c7_iti_accuracy.txt 2>&1; cat gpurun_out/c7_iti_accuracy.txt | tail -20
timeout 900 python bench.py --steps 2 --warmup 3 --host-resident 0 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','factored_root','stages']})
print(d['roofline']['other_kernels_ms_per_step'], d['roofline']['hbm_kernels'])
PY
tail -3 gpurun_out/c7_bench.err
