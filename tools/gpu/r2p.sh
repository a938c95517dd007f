#!/bin/bash
# round-2 GPU batch: full suite with the stencil wavefront chosen automatically, configs table
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/r2p_pytest.log | cut -c1-250
timeout 600 python tools/bench_configs.py --skip-cpu > gpurun_out/r2p_configs.jsonl 2> gpurun_out/r2p_configs.err; echo "configs rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2p_configs.jsonl'):
    d=json.loads(l); print(d['config'], d['matrix'][:12], d['method'], d['precond'], 'ms/iter', round(d['gpu_ms_per_iter'],4), 'eager+prof', round(d['gpu_ms_per_iter_profiled_eager'],4), 'spmv', round(d['spmv_ms_per_iter'],4),'trsv', round(d['sptrsv_ms_per_iter'],4), 'vec', round(d['vector_ms_per_iter'],4), 'launches/it', d['launches_per_iter'])
PY
