"""Sweep the SpMV launch parameters on a device-generated HPCG-n matrix (run on the GPU box):
   python tools/tune_spmv.py [n] > gpurun_out/tune_spmv.txt
Prints achieved GB/s on the algorithmic bytes (12 nnz + rp (n+1) + 16 n) per configuration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = 10
with capi.Context(0) as ctx:
    A = ctx.generate_hpcg(n)
    inf = A.info()
    x, y = ctx.alloc(inf["n_rows"]), ctx.alloc(inf["n_rows"])
    ctx.call("bis_init_vector", x, 1.0, inf["n_rows"])
    nbytes = A.spmv_bytes()
    print(f"HPCG-{n}: rows {inf['n_rows']} nnz {inf['nnz']} rp_bytes {inf['rp_bytes']} bytes/spmv {nbytes/1e9:.3f} GB")

    w = ctx.alloc(inf["n_rows"])
    ctx.call("bis_init_vector", w, 0.5, inf["n_rows"])
    kind = {"v": "spmv"}

    def launch():
        if kind["v"] == "spmv":
            ctx.call("bis_spmv", A.h, x, y)
        elif kind["v"] == "dot_self":      # CG: w is x itself
            ctx.call("bis_spmv_dot", A.h, x, y, x, 40, -1)
        elif kind["v"] == "dot":
            ctx.call("bis_spmv_dot", A.h, x, y, w, 40, 41)
        elif kind["v"] == "resid":
            ctx.call("bis_spmv_residual", A.h, x, w, y, None, 40)
        elif kind["v"] == "jacobi":
            ctx.call("bis_spmv_jacobi", A.h, w, w, x, y)
        elif kind["v"] == "sub":
            ctx.call("bis_spmv_sub", A.h, x, w, y)

    def run(label, **opts):
        for k, v in opts.items():
            ctx.set_option(k, v)
        for _ in range(2):
            launch()
        ctx.sync()
        ctx.timer_start()
        for _ in range(reps):
            launch()
        ms = ctx.timer_stop() / reps
        print(f"{label:55s} {ms:8.3f} ms  {nbytes/ms/1e6:8.1f} GB/s", flush=True)
        for k in opts:
            ctx.set_option(k, 0)

    for k in ("spmv", "sub", "jacobi", "resid", "dot", "dot_self"):
        kind["v"] = k
        run(f"{k}: default (auto)")
