#!/bin/bash
# round-2 final measurement batch (1 GPU): smoke, bench (both arms), launch list + full capture of the top kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2q_smoke.log | cut -c1-300
python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2q_bench.json; tail -n 3 gpurun_out/r2q_bench.err
python bench.py --impl reference > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/r2q_bench_ref.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/r2q_bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench512.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solve > gpurun_out/r2q_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_bench512.csv "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solve" > gpurun_out/r02_launches_bench512_summary.txt 2>&1; head -n 30 gpurun_out/r02_launches_bench512_summary.txt | cut -c1-200
