"""Run a few forward triangular solves on HPCG-n's strict lower factor (profiling target):
   python tools/run_trsv.py n reps [key=value ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi, host  # noqa: E402

n, reps = int(sys.argv[1]), int(sys.argv[2])
rp, col, val = host.matrix(f"HPCG-{n}")
f = host.factor(rp, col, val, "sgs")
with capi.Context(0) as ctx:
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    L = ctx.upload_triangular(f.l_rp, f.l_col, f.l_val, upper=False)
    N = len(rp) - 1
    D, b, x = ctx.upload(f.A_D), ctx.upload(np.ones(N)), ctx.alloc(N)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        ctx.call("bis_sptrsv", L.h, x, D, b)
    ms = ctx.timer_stop() / reps
    inf = L.info()
    print(f"HPCG-{n} forward solve {sys.argv[3:]}: {ms:.3f} ms, {inf['n_levels']} levels, {1e3*ms/inf['n_levels']:.2f} us/level, "
          f"{(12*inf['nnz'] + 44*N)/ms/1e6:.0f} GB/s")
