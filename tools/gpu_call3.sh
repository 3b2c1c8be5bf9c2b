#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x -s -k "config3 or stepwise or lu_solve or sharded_driver" > gpurun_out/c3_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest_a.log
grep -E "config 3|passed|failed|rc=" gpurun_out/c3_pytest_a.log | tail -8
rm -f gpurun_out/c3_bench_lu.txt
for cfg in "19200 1 0" "4800 1 0" "1000 512 600"; do
  timeout 300 python tools/bench_lu.py $cfg >> gpurun_out/c3_bench_lu.txt 2>&1
done
grep -v "^\[W" gpurun_out/c3_bench_lu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
tail -5 gpurun_out/c3_pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --host-resident 0 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
cat gpurun_out/c3_bench.json; tail -5 gpurun_out/c3_bench.err
