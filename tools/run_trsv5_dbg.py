"""Perf ablations of the stencil-wavefront solve (needs a -DBIS_PERF_DEBUG build): python tools/run_trsv5_dbg.py n"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

n = int(sys.argv[1])
with capi.Context(0) as ctx:
    A = ctx.generate_hpcg(n)
    N = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(N)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    b, x = ctx.upload(np.ones(N)), ctx.alloc(N)
    # never combine 16 (no w stores) without 8 (no requests): readers would wait for ever; never use 1 alone
    for dbg in [0, 32, 2, 4, 8, 8 | 16, 8 | 32, 8 | 2, 8 | 16 | 32 | 2 | 4]:
        try:
            ctx.set_option("wave_debug", dbg)
        except Exception as e:  # not a debug build
            print("no debug build:", e)
            break
        ctx.call("bis_sptrsv", L.h, x, D, b)
        ctx.sync()
        ctx.timer_start()
        for _ in range(5):
            ctx.call("bis_sptrsv", L.h, x, D, b)
        ms = ctx.timer_stop() / 5
        steps = n + 62 + 8
        print(f"HPCG-{n} forward, wave_debug={dbg:2d}: {ms:.3f} ms" + (f"  ({1e3 * ms / steps:.3f} us per step, no dependencies)" if dbg & 8 else ""), flush=True)
