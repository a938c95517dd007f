#!/bin/bash
# round-2 GPU batch A: full GPU test suite, bench line, SpMV traversal experiment, configs table
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt
timeout 1700 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"
tail -n 15 gpurun_out/r2a_pytest.log
for mb in 40 100000 20 80; do
  timeout 200 python tools/run_spmv.py 512 10 kind=dot spmv_l2_mb=$mb >> gpurun_out/r2a_spmv_l2.txt 2>&1
done
cat gpurun_out/r2a_spmv_l2.txt
timeout 900 python bench.py --no-cpu-baseline --write-parity-golden gpurun_out/bench_parity.json > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -n 5 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json
timeout 600 python tools/bench_configs.py --skip-cpu > gpurun_out/r2a_configs.jsonl 2> gpurun_out/r2a_configs.err
echo "configs rc=$?"; cat gpurun_out/r2a_configs.jsonl | cut -c1-600
