"""Diagnose a wrong row of the stencil-wavefront forward solve on an HPCG grid (all off-diagonals -1, diagonal 26):
   python tools/run_trsv5_diag.py 224x225x8 [tries]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1].split("x"))
tries = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with capi.Context(0) as ctx:
    A = ctx.generate_hpcg(nx, ny, nz)
    N = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(N)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    bh = np.random.default_rng(3).uniform(-1.0, 1.0, N)
    b, x = ctx.upload(bh), ctx.alloc(N)
    ctx.set_option("trsv_variant", 3)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
    want = ctx.download(x, N)
    ctx.set_option("trsv_variant", 5)
    found = 0
    for t in range(tries):
        ctx.call("bis_sptrsv", L.h, x, D, b)
        ctx.sync()
        got = ctx.download(x, N)
        bad = np.nonzero(got != want)[0]
        if bad.size == 0:
            continue
        found += 1
        r = int(bad[0])
        xx, yy, zz = r % nx, (r // nx) % ny, r // (nx * ny)
        e = (want[r] - got[r]) * 26.0          # error of the numerator b - sum (got = (b - sum') / 26)
        print(f"try {t}: {bad.size} rows differ; first row {r} = (x={xx}, y={yy}, z={zz}); want {want[r]!r} got {got[r]!r}; numerator error {e!r}")
        print(f"   b[r] = {bh[r]!r}, b[r-1] = {bh[r-1]!r}, b[r+1] = {bh[min(r+1, N-1)]!r}")
        for dz in (-1, 0):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if (dz, dy, dx) >= (0, 0, 0):
                        continue
                    X, Y, Z = xx + dx, yy + dy, zz + dz
                    if 0 <= X < nx and 0 <= Y < ny and 0 <= Z < nz:
                        c = (Z * ny + Y) * nx + X
                        print(f"   neighbour ({dx:+d},{dy:+d},{dz:+d}) row {c}: x = {want[c]!r}  (got {got[c]!r})  ratio e/x = {e / want[c] if want[c] else float('nan'):.6f}")
        # second differing row for context
        print(f"   next differing rows: {bad[1:8]}")
        if found >= 3:
            break
    print(f"{found} wrong solves in {t + 1} tries")
