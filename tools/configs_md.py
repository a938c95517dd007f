"""Render gpurun_out/configs.jsonl (tools/bench_configs.py) as the markdown table kept under profiles/:
   python tools/configs_md.py gpurun_out/configs.jsonl > profiles/r01_configs.md"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
cores = next((r["cpu_cores"] for r in rows if "cpu_cores" in r), "?")
print("# BASELINE configs on one B200 vs the unmodified reference on the box's host cores (round 1)\n")
print(f"`python tools/bench_configs.py` (20 timed iterations after 3 warm-up, CUDA events; reference: same iteration "
      f"count, its own stopwatches, {cores} OpenMP threads).\n")
print("| config | matrix / solver | GPU ms/iter | SpMV ms/iter | SpTRSV ms/iter (per sweep) | vector ms/iter | SpMV GB/s "
      "(algorithmic) | GPU setup s | CPU ms/iter | speed-up | max |r_k-r_k^ref|/r0 | max rel |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    name = r["matrix"].split(",")[0]
    solver = f"-{r['method']}" + ("" if r["precond"] == "none" else f" -p {r['precond']}")
    cpu = f"{r['cpu_ms_per_iter']:.2f}" if "cpu_ms_per_iter" in r else "–"
    sp = f"{r['speedup_vs_cpu']:.1f}" if "speedup_vs_cpu" in r else "–"
    d1 = f"{r['history_max_abs_diff_over_r0']:.1e}" if "history_max_abs_diff_over_r0" in r else "–"
    d2 = f"{r['history_max_rel_diff']:.1e}" if "history_max_rel_diff" in r else "–"
    sw = f" ({r['sptrsv_ms_per_sweep']:.3f})" if "sptrsv_ms_per_sweep" in r else ""
    print(f"| {r['config']} | {name} {solver} | {r['gpu_ms_per_iter']:.3f} | {r['spmv_ms_per_iter']:.3f} | "
          f"{r['sptrsv_ms_per_iter']:.3f}{sw} | {r['vector_ms_per_iter']:.3f} | {r.get('spmv_gbs', 0):.0f} | "
          f"{r['gpu_preprocessing_s']:.2f} | {cpu} | {sp} | {d1} | {d2} |")
print("\nNotes: config 4 (Anderson, ILU(0) without pivoting on an indefinite matrix) is numerically unstable in the "
      "reference itself (preconditioned residuals ~1e60, SURVEY.md §7); 4s is the same matrix with the Jacobi "
      "preconditioner (division by near-zero diagonal entries: compare the relative column). The reference cannot hold "
      "HPCG-512 (32-bit nnz), so configs 5/5c have no CPU column here; bench.py scales an HPCG-256 run by the row ratio. "
      "\"GPU setup\" is preprocessing() of the host stack: matrix generation on the device, and for triangular methods the "
      "split / ILU(0) / level analysis.")
