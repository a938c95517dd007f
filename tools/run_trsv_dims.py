"""Forward triangular solves on the strict lower factor of an HPCG nx x ny x nz grid (device-side split):
   python tools/run_trsv_dims.py nx ny nz reps [key=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

nx, ny, nz, reps = (int(v) for v in sys.argv[1:5])
with capi.Context(0) as ctx:
    for kv in sys.argv[5:]:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    A = ctx.generate_hpcg(nx, ny, nz)
    L, U = ctx.split_triangular(A)
    n = nx * ny * nz
    D, b, x = ctx.alloc(n), ctx.alloc(n), ctx.alloc(n)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    ctx.call("bis_init_vector", b, 1.0, n)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        ctx.call("bis_sptrsv", L.h, x, D, b)
    ms = ctx.timer_stop() / reps
    inf = L.info()
    print(f"HPCG {nx}x{ny}x{nz} forward solve {sys.argv[5:]}: {ms:.3f} ms, {inf['n_levels']} levels, "
          f"{1e3*ms/inf['n_levels']:.2f} us/level, chain solves {ctx.info()['chain_solves']}", flush=True)
