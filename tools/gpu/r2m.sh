#!/bin/bash
# multi-GPU batch: partition-invariant reductions (bit-identical histories), bench line with parity block
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
for n in 48 96; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py $n > gpurun_out/r2m_dist${N}_hpcg$n.log 2>&1
  echo "dist_check $n rc=$?"; grep -E "DIST_CHECK|BIT-IDENTICAL|differ|transport|unstructured|Error|error" gpurun_out/r2m_dist${N}_hpcg$n.log | cut -c1-200
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
echo "bench rc=$?"; tail -n 3 gpurun_out/r2m_bench_n$N.err; python - <<PY
import json
d=json.load(open("gpurun_out/r2m_bench_n$N.json"))
print({k:d[k] for k in ("value","n_gpus","parity","dist_wait")})
print("e2e",d["e2e"]["value"],"spmv ms",d["roofline"]["ms_per_launch"],"frac",d["roofline"]["frac"], "bi", d["also"])
PY
