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

    for lanes in (4, 8, 16):
        run(f"vector lanes={lanes}", spmv_variant=1, spmv_lanes=lanes)
    for blocked in (0, 1):
        for rows, stages, kb, mult in ((256, 2, 0, 1), (256, 2, 0, 2), (256, 2, 0, 4), (128, 2, 100, 2), (128, 2, 100, 4),
                                       (128, 2, 100, 8), (64, 2, 50, 2), (64, 2, 50, 4), (128, 4, 0, 4), (128, 4, 0, 8),
                                       (64, 2, 70, 4), (64, 3, 70, 4)):
            run(f"tma rows={rows} stages<={stages} smem_kb={kb or 200} mult={mult} blocked={blocked}", spmv_rows=rows,
                spmv_stages=stages, spmv_smem_kb=kb, spmv_blocked=blocked, spmv_mult=mult)
