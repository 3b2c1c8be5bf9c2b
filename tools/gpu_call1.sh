#!/bin/bash
# Round-2 GPU call 1 (one GPU): GPU test suite with the new block-column LU kernel, LU micro-bench, GEMM lab,
# ncu evidence for the product GEMM and the LU kernels.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
for cfg in "1000 512 600" "1200 64 2400" "4800 8 9600" "19200 1 0" "19200 1 4800" "512 256 64" "196 4096 57"; do
  timeout 300 python tools/bench_lu.py $cfg >> gpurun_out/c1_bench_lu.txt 2>&1
done
cat gpurun_out/c1_bench_lu.txt
timeout 400 ./tools/gemm_lab > gpurun_out/r02_gemm_lab.txt 2>&1
cat gpurun_out/r02_gemm_lab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel_hoist -s 2 -c 1 \
    -o gpurun_out/r02_gemm_hoist_8192 python tools/one_gemm.py 8192 8192 8192 > gpurun_out/r02_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel_hoist -s 2 -c 1 \
    -o gpurun_out/r02_gemm_hoist_k128 python tools/one_gemm.py 15360 15360 128 1.0 > gpurun_out/r02_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blockcol_kernel -s 8 -c 2 \
    -o gpurun_out/r02_blockcol_cluster python tools/bench_lu.py 1000 256 0 1 > gpurun_out/r02_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blockcol_kernel -s 150 -c 2 \
    -o gpurun_out/r02_blockcol_coop python tools/bench_lu.py 19200 1 0 1 > gpurun_out/r02_ncu4.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
cat gpurun_out/c1_bench.json
ls -la gpurun_out
