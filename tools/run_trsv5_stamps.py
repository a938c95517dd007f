"""Per-plane start / end times of one stencil-wavefront solve (needs a -DBIS_PERF_DEBUG build):
   python tools/run_trsv5_stamps.py n"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

n = int(sys.argv[1])
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 0
path = "/tmp/wave_stamps.bin"
os.environ["BIS_WAVE_STAMPS"] = path
with capi.Context(0) as ctx:
    A = ctx.generate_hpcg(n)
    N = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(N)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    b, x = ctx.upload(np.ones(N)), ctx.alloc(N)
    ctx.set_option("trsv_variant", 5)
    if extra:
        ctx.set_option("wave_debug", extra)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
    ctx.set_option("wave_debug", 64 | extra)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
raw = np.fromfile(path, dtype=np.uint64).astype(np.int64)
acc = raw[-64:]
t = raw[:-64].reshape(-1, 2)
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
print(f"HPCG-{n}: {t.shape[0]} planes; plane 0 runs {end[0] - start[0]:.1f} us; last plane ends at {end[-1]:.1f} us")
d_end = np.diff(end)
print(f"end-to-end lag between consecutive planes: median {np.median(d_end):.2f} us, mean {d_end.mean():.2f}, min {d_end.min():.2f}, max {d_end.max():.2f}")
print(f"duration of a plane (start to end): median {np.median(end - start):.1f} us, first {end[0]-start[0]:.1f}, last {end[-1]-start[-1]:.1f}")
for cl in (8,):
    lag = np.diff(end)
    by_rank = [np.median(lag[(np.arange(1, t.shape[0]) % cl) == r]) for r in range(cl)]
    print(f"end-to-end lag by plane index mod {cl}: " + ", ".join(f"{v:.2f}" for v in by_rank) + " us")
    dst = np.diff(start)
    by_rank = [np.median(dst[(np.arange(1, t.shape[0]) % cl) == r]) for r in range(cl)]
    print(f"start-to-start lag by plane index mod {cl}: " + ", ".join(f"{v:.2f}" for v in by_rank) + " us")
for z in list(range(0, 6)) + [t.shape[0] // 2, t.shape[0] - 1]:
    print(f"  plane {z:4d}: start {start[z]:9.1f} us, end {end[z]:9.1f} us")

for slot, zname in enumerate(("1", "8", "33", "nz/2", "nz/2+4")):
    for ph in (0, 1):
        a = acc[(slot * 2 + ph) * 6: (slot * 2 + ph) * 6 + 6]
        if a[4] > 0:
            names = ("solve", "barrier", "prepare", "barrier") if ph == 0 else ("prepare", "barrier", "refill+solve", "barrier")
            print(f"  plane {zname:>6s}, block 1, warp {ph}: per step pair " + ", ".join(f"{nm} {v / a[4]:.0f}" for nm, v in zip(names, a[:4]))
                  + f" cycles; total {a[:4].sum() / a[4]:.0f}; {a[5]} polls in the whole plane ({a[4]} mid-plane pairs)")
