"""All five BASELINE.json configs on one B200, next to the unmodified reference on the host cores:
   python tools/bench_configs.py [--steps 20] [--skip-cpu] > gpurun_out/configs.jsonl
Per config: GPU ms/iteration (CUDA events, state resident), per-family kernel time (SpMV / SpTRSV /
vector), SpMV GB/s on the algorithmic bytes, the reference's ms/iteration from its own stopwatches for the
same number of iterations, and the parity of the two residual histories over those iterations."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi, host  # noqa: E402

CONFIGS = [
    ("1", "HPCG-128", "cg", "none", 10),
    ("2a", "HPCG-128", "j", "none", 10),
    ("2b", "HPCG-128", "sgs", "none", 10),
    ("3", "HPCG-256", "cg", "sgs", 10),
    ("4", "Anderson,Lx=100,Ly=100,Lz=50,ranpot=5.0", "gm", "ilu0", 10),
    ("4s", "Anderson,Lx=100,Ly=100,Lz=50,ranpot=5.0", "gm", "j", 10),
    ("5", "HPCG-512", "bi", "j", 10),
    ("5c", "HPCG-512", "cg", "j", 10),
]

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--skip-cpu", action="store_true")
ap.add_argument("--only", default="")
args = ap.parse_args()
K, W = args.steps, 3
peak = 6540.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

with capi.Context(0) as ctx:
    for tag, name, method, pre, rl in CONFIGS:
        if args.only and tag not in args.only.split(","):
            continue
        out = {"config": tag, "matrix": name, "method": method, "precond": pre}
        # per-launch event pairs need the launches issued one by one: graph replay off for this pass (with it on, the
        # iterations whose graph had been captured during the warm-up replayed without brackets and the per-family
        # sums of the first two rounds' tables came out at half their value for the methods with two graphs)
        ctx.set_option("graph", 0)
        sess = host.BenchSession(ctx, name, method, pre, rl)
        t0 = time.time()
        info = sess.prepare(W)
        out["gpu_preprocessing_s"] = time.time() - t0
        ctx.profile_enable(True)
        r = sess.run(K)
        fam = {f: ctx.profile_read(f) for f in ("spmv", "sptrsv", "vector")}
        ctx.profile_enable(False)
        ctx.set_option("graph", 1)
        sess.close()
        # the number that counts: profiling off, iteration bodies replayed as CUDA graphs
        sess = host.BenchSession(ctx, name, method, pre, rl)
        sess.prepare(W + 4)
        rg = sess.run(K)
        sess.close()
        out["gpu_ms_per_iter_profiled_eager"] = r["device_ms"] / K
        out.update({"rows": info["n_rows"], "nnz": info["nnz"], "gpu_ms_per_iter": rg["device_ms"] / K,
                    "launches_per_iter": r["launches"] / K,
                    "spmv_ms_per_iter": fam["spmv"][0] / K, "sptrsv_ms_per_iter": fam["sptrsv"][0] / K,
                    "vector_ms_per_iter": fam["vector"][0] / K})
        if fam["spmv"][1]:
            b = 12 * info["nnz"] + info["rp_bytes"] * (info["n_rows"] + 1) + 16 * info["n_rows"]
            ms = fam["spmv"][0] / fam["spmv"][1]
            out["spmv_gbs"] = b / ms / 1e6
            out["spmv_frac_of_measured_peak"] = out["spmv_gbs"] / peak
        if fam["sptrsv"][1]:
            out["sptrsv_ms_per_sweep"] = fam["sptrsv"][0] / fam["sptrsv"][1]
        # parity + CPU baseline over the same W+K iterations (reference cannot hold HPCG-512: F5)
        if not args.skip_cpu and "512" not in name:
            from oracle import refshim
            lib = refshim.load()
            cores = len(os.sched_getaffinity(0))
            lib.ref_omp_set_threads(cores)
            rp, col, val = host.matrix(name)
            g = host.solve(ctx, method, pre, crs=(rp, col, val), restart_len=rl, max_iters=W + K, tol=1e-300, want_x=False)
            lib.ref_set_max_iters(W + K)
            t0 = time.time()
            c = refshim.solve(rp, col, val, method, pre, restart_len=rl)
            out["cpu_wall_s"] = time.time() - t0
            lib.ref_set_max_iters(0)
            its = max(c.iter_count, 1)
            out.update({"cpu_cores": cores, "cpu_ms_per_iter": 1e3 * c.iterate_time / its,
                        "cpu_spmv_ms_per_iter": 1e3 * c.spmv_time / its,
                        "cpu_precond_ms_per_iter": 1e3 * c.precond_time / its})
            k = min(g.history.size, c.history.size)
            with np.errstate(invalid="ignore", divide="ignore"):
                out["history_max_abs_diff_over_r0"] = float(np.nanmax(np.abs(g.history[:k] - c.history[:k])) / c.history[0])
                out["history_max_rel_diff"] = float(np.nanmax(np.abs(g.history[:k] - c.history[:k]) / np.abs(c.history[:k])))
            out["history_entries_compared"] = int(k)
            out["speedup_vs_cpu"] = out["cpu_ms_per_iter"] / out["gpu_ms_per_iter"]
        print(json.dumps(out), flush=True)
