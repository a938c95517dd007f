"""How often the stencil-wavefront forward solve differs from the dataflow solve: python tools/run_trsv5_count.py 224x225x8 tries"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1].split("x"))
tries = int(sys.argv[2])
with capi.Context(0) as ctx:
    A = ctx.generate_hpcg(nx, ny, nz)
    N = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(N)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    b, x = ctx.upload(np.random.default_rng(3).uniform(-1.0, 1.0, N)), ctx.alloc(N)
    want = {}
    ctx.set_option("trsv_variant", 3)
    for T, fn in ((L, "bis_sptrsv"), (U, "bis_bsptrsv")):
        ctx.call(fn, T.h, x, D, b)
        ctx.sync()
        want[fn] = ctx.download(x, N)
    ctx.set_option("trsv_variant", 5)
    bad = {"bis_sptrsv": 0, "bis_bsptrsv": 0}
    for t in range(tries):
        for T, fn in ((L, "bis_sptrsv"), (U, "bis_bsptrsv")):
            ctx.call(fn, T.h, x, D, b)
            ctx.sync()
            bad[fn] += int(not np.array_equal(ctx.download(x, N), want[fn]))
    print(f"{sys.argv[1]}: wrong forward solves {bad['bis_sptrsv']} / {tries}, wrong backward solves {bad['bis_bsptrsv']} / {tries}")
