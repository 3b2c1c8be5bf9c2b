#!/bin/bash
# 1 GPU: parallel laswp + priority look-ahead column: correctness, LU micro-benchmarks (new vs serial laswp),
# ncu source-level captures of the block-column kernel (n=19200, cooperative) and the cluster panel kernel (n=1000 x512).
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_stages.py tests/test_gpu_e2e.py -m gpu -q -x > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/c18_pytest.log
for cfg in "1000 512 601" "1200 64 2401" "4800 8 9601" "19200 1 0" "19200 1 38401"; do
  for mode in block serial; do
    echo "== $cfg laswp=$mode"; HPS_LASWP=$mode timeout 300 python tools/bench_lu.py $cfg 2 2>&1 | tail -4
  done
done > gpurun_out/c18_bench_lu.txt 2>&1
grep -E "==|iter 2|per category" gpurun_out/c18_bench_lu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blockcol_kernel -s 4 -c 1 \
    -o gpurun_out/c18_blockcol python tools/bench_lu.py 19200 1 0 0 > gpurun_out/c18_ncu_blockcol.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 4 -c 1 \
    -o gpurun_out/c18_panel python tools/bench_lu.py 1000 512 0 0 > gpurun_out/c18_ncu_panel.log 2>&1; echo "ncu2 rc=$?"
timeout 600 python bench.py --steps 2 --warmup 2 --factored 0 --host-resident 0 > gpurun_out/c18_bench_n1.json 2> gpurun_out/c18_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c18_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','max_rel_error_vs_analytic_solution']}, d['stages'], d['e2e'])
r=d['roofline']; print({k:r.get(k) for k in ['achieved','peak','frac','gemm_ms_per_step','other_kernels_ms_per_step']})
PY
ls -la gpurun_out | head -30
