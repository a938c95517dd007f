"""One configuration of the stencil-wavefront solve (profiling target): python tools/run_trsv5_one.py n wave_debug reps"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

n, dbg, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
with capi.Context(0) as ctx:
    A = ctx.generate_hpcg(n)
    N = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(N)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    b, x = ctx.upload(np.ones(N)), ctx.alloc(N)
    ctx.set_option("trsv_variant", 5)
    if dbg:
        ctx.set_option("wave_debug", dbg)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        ctx.call("bis_sptrsv", L.h, x, D, b)
    print(f"HPCG-{n} wave_debug={dbg}: {ctx.timer_stop() / reps:.3f} ms")
