"""One forward + one backward stencil-wavefront solve on an nx x ny x nz HPCG grid (target for compute-sanitizer):
   python tools/run_trsv5_grid.py 224x225x8 [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1].split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dbg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
with capi.Context(0) as ctx:
    A = ctx.generate_hpcg(nx, ny, nz)
    N = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(N)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    b, x = ctx.upload(np.random.default_rng(3).uniform(-1.0, 1.0, N)), ctx.alloc(N)
    out = {}
    if dbg:
        ctx.set_option("wave_debug", dbg)
    for variant in (5, 3):
        ctx.set_option("trsv_variant", variant)
        bad = 0
        for _ in range(reps):
            ctx.call("bis_sptrsv", L.h, x, D, b)
            ctx.sync()
            f = ctx.download(x, N)
            ctx.call("bis_bsptrsv", U.h, x, D, b)
            ctx.sync()
            g = ctx.download(x, N)
            if variant == 3:
                out[3] = (f, g)
            else:
                out.setdefault(5, []).append((f, g))
    nbad = sum(1 for f, g in out[5] if not (np.array_equal(f, out[3][0]) and np.array_equal(g, out[3][1])))
    print(f"{sys.argv[1]} wave_debug={dbg}: {nbad} of {reps} wavefront solve pairs differ from the dataflow solve")
