#!/bin/bash
# final single-GPU batch: full suite, smoke, both bench arms, configs table
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r3z_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r3z_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3z_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r3z_smoke.log | cut -c1-300
python bench.py > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/r3z_bench.json
python bench.py --impl reference > gpurun_out/r3z_bench_ref.json 2> gpurun_out/r3z_bench_ref.err; echo "ref rc=$?"; cut -c1-120 gpurun_out/r3z_bench_ref.json
timeout 600 python tools/bench_configs.py --skip-cpu > gpurun_out/r3z_configs.jsonl 2> gpurun_out/r3z_configs.err; echo "configs rc=$?"
