#!/bin/bash
# SpMV (dot-fused, HPCG-512) tile-shape sweep
cd "${GRAFT_REPO_ROOT:-.}"
for opts in "" "win_rows=64" "win_rows=64 spmv_smem_kb=75" "win_rows=64 spmv_smem_kb=100 spmv_stages=8" "win_rows=96" "win_rows=96 spmv_smem_kb=75" "win_rows=192 spmv_smem_kb=220" "win_rows=256 spmv_smem_kb=220" "spmv_smem_kb=150" "spmv_smem_kb=220 spmv_stages=4" "win_rows=160 spmv_smem_kb=113"; do
  timeout 120 python tools/run_spmv.py 512 5 kind=dot $opts 2>&1 | tail -n 1
done
