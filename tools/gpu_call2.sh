#!/bin/bash
# Round-2 GPU call 2 (one GPU): GPU test suite, LU micro-bench, ncu of the block-column kernel, bench N=1.
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -15 gpurun_out/c2_pytest.log
rm -f gpurun_out/c2_bench_lu.txt
for cfg in "1000 512 600" "1200 64 2400" "4800 8 9600" "19200 1 0" "19200 1 4800" "512 256 64" "196 4096 57"; do
  timeout 300 python tools/bench_lu.py $cfg >> gpurun_out/c2_bench_lu.txt 2>&1
done
cat gpurun_out/c2_bench_lu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blockcol_kernel -s 8 -c 2 \
    -o gpurun_out/r02_blockcol_cluster python tools/bench_lu.py 1000 256 0 1 > gpurun_out/r02_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blockcol_kernel -s 150 -c 2 \
    -o gpurun_out/r02_blockcol_coop python tools/bench_lu.py 19200 1 0 1 > gpurun_out/r02_ncu4.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
cat gpurun_out/c2_bench.json; tail -5 gpurun_out/c2_bench.err
ls -la gpurun_out
