"""Set-up time of a generated matrix: generation, then the first SpMV (which builds the tile format and the order
table), wall clock around synchronised calls: python tools/run_setup_time.py n"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

n = int(sys.argv[1])
with capi.Context(0) as ctx:
    for rep in range(3):
        ctx.sync()
        t0 = time.time()
        A = ctx.generate_hpcg(n)
        ctx.sync()
        t1 = time.time()
        N = A.info()["n_rows"]
        x, y = ctx.alloc(N), ctx.alloc(N)
        ctx.call("bis_init_vector", x, 1.0, N)
        ctx.sync()
        t2 = time.time()
        ctx.call("bis_spmv", A.h, x, y)
        ctx.sync()
        t3 = time.time()
        ctx.call("bis_spmv", A.h, x, y)
        ctx.sync()
        t4 = time.time()
        print(f"HPCG-{n}: generation {1e3 * (t1 - t0):.1f} ms, first SpMV (tile format + order table + launch) {1e3 * (t3 - t2):.1f} ms, second SpMV {1e3 * (t4 - t3):.1f} ms")
        A.free()
        ctx.free(x)
        ctx.free(y)
