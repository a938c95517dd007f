"""Read-only and copy streaming bandwidth of this library's element-wise kernels (context for the SpMV roofline):
   python tools/run_readbw.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

n = 1 << 29          # 4.3 GB per vector
with capi.Context(0) as ctx:
    a, b, o = ctx.alloc(n), ctx.alloc(n), ctx.alloc(n)
    ctx.call("bis_init_vector", a, 1.0, n)
    ctx.call("bis_init_vector", b, 0.5, n)
    for name, fn, nbytes in (
        ("dot (2 reads)", lambda: ctx.call("bis_dot_to_slot", a, b, n, 50), 16 * n),
        ("norm^2 (1 read)", lambda: ctx.call("bis_dot_to_slot", a, a, n, 50), 8 * n),
        ("copy (1 read + 1 write)", lambda: ctx.call("bis_copy_vector", o, a, n), 16 * n),
        ("sum (2 reads + 1 write)", lambda: ctx.call("bis_sum_vectors", o, a, b, n, 1.0), 24 * n),
    ):
        fn()
        ctx.sync()
        ctx.timer_start()
        for _ in range(10):
            fn()
        ms = ctx.timer_stop() / 10
        print(f"{name:28s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.1f} GB/s")
