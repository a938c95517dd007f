#!/bin/bash
# final multi-GPU batch: dist_check at HPCG-96 + bench at N GPUs, then (if asked) bench at N/2
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 96 > gpurun_out/r2s_dist${N}_hpcg96.log 2>&1
echo "dist_check rc=$?"; grep -E "DIST_CHECK|BIT-IDENTICAL|differ|transport|unstructured|Error|error" gpurun_out/r2s_dist${N}_hpcg96.log | cut -c1-200
for M in $N $2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $M > gpurun_out/r2s_bench_n$M.json 2> gpurun_out/r2s_bench_n$M.err
  echo "bench N=$M rc=$?"; tail -n 2 gpurun_out/r2s_bench_n$M.err; python - <<PY
import json
d=json.load(open("gpurun_out/r2s_bench_n$M.json"))
print({k:d[k] for k in ("value","n_gpus","parity","dist_wait")})
print("e2e",d["e2e"]["value"],"spmv ms",d["roofline"]["ms_per_launch"],"frac",d["roofline"]["frac"], "vec", d["vector_kernels_ms_per_iter"], "bi", d["also"])
PY
done
