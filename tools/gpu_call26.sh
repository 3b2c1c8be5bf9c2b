#!/bin/bash
# 1 GPU: speculative block columns (in-block pivoting) in the adaptive merges: GPU suite, configs 5 and 4, default bench,
# ncu captures of the new chain kernels.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c26_pytest.log
for spec in 1 0; do
  HPS_LU_SPEC=$spec timeout 600 python tools/run_config5.py --repeat 2 --prof --load-tree tools/data/config5_tree_p10_tol1e-3.npy > gpurun_out/c26_config5_spec$spec.log 2>&1
  echo "config5 spec=$spec rc=$?"; tail -3 gpurun_out/c26_config5_spec$spec.log | cut -c1-900
done
timeout 400 python tools/run_config4.py --p 10 --tol 1e-5 --repeat 2 > gpurun_out/c26_config4.log 2>&1; tail -2 gpurun_out/c26_config4.log | cut -c1-600
timeout 600 python bench.py > gpurun_out/c26_bench_n1.json 2> gpurun_out/c26_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c26_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution','same_config_sample','e2e_host_resident']}, d['stages'], d['e2e'], d['factored_root'], d['cpu_baseline'])
r=d['roofline']; print({k:r.get(k) for k in ['achieved','peak','frac','gemm_ms_per_step','other_kernels_ms_per_step','hbm_kernels']})
PY
HPS_LU_FORCE_SPEC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"diagblk_kernel|trtri_pair_kernel|spec_commit_kernel" -s 6 -c 3 \
    -o gpurun_out/c26_chain_kernels python tools/bench_lu.py 19200 1 0 0 > gpurun_out/c26_ncu_chain.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 3000 --csv --log-file gpurun_out/c26_launches.csv python bench.py --steps 1 --warmup 1 --factored 0 --host-resident 0 > gpurun_out/c26_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
