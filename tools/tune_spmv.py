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

    def run(label, **opts):
        for k, v in opts.items():
            ctx.set_option(k, v)
        for _ in range(2):
            ctx.call("bis_spmv", A.h, x, y)
        ctx.sync()
        ctx.timer_start()
        for _ in range(reps):
            ctx.call("bis_spmv", A.h, x, y)
        ms = ctx.timer_stop() / reps
        print(f"{label:55s} {ms:8.3f} ms  {nbytes/ms/1e6:8.1f} GB/s", flush=True)
        for k in opts:
            ctx.set_option(k, 0)

    run("default (auto)")
    for stages, kb in ((2, 0), (3, 0), (2, 75), (3, 150), (4, 200)):
        run(f"win (variant 3) stages<={stages} smem_kb={kb or 110}", spmv_variant=3, spmv_stages=stages, spmv_smem_kb=kb)
    run("win, debug=1 no compute (supply ceiling; result invalid)", spmv_variant=3, spmv_debug=1)
    run("win, debug=2 no x copies (result invalid)", spmv_variant=3, spmv_debug=2)
    run("tma (variant 2) default", spmv_variant=2)
