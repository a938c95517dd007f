#!/bin/bash
# round-2 GPU batch B: stencil-wavefront triangular solve -- correctness vs the dataflow solve, timings, suite
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python tools/run_trsv5.py check > gpurun_out/r2b_check.txt 2>&1; echo "check rc=$?"; cat gpurun_out/r2b_check.txt | tail -n 20
for n in 128 256; do timeout 300 python tools/run_trsv5.py time $n 10 >> gpurun_out/r2b_time.txt 2>&1; done; cat gpurun_out/r2b_time.txt
timeout 1700 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/r2b_pytest.log
timeout 600 python tools/bench_configs.py --skip-cpu --only 2b,3,4 > gpurun_out/r2b_configs.jsonl 2> gpurun_out/r2b_configs.err
echo "configs rc=$?"; cut -c1-700 gpurun_out/r2b_configs.jsonl; tail -n 3 gpurun_out/r2b_configs.err
