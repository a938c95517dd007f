"""Variant 5 of the triangular solve (stencil wavefront) against the dataflow solve (variant 3), bit for bit, on
device-generated factors, plus timings:
   python tools/run_trsv5.py check            # correctness on a set of grids (forward, backward, in place, SGS apply)
   python tools/run_trsv5.py time n [reps]    # per-sweep time of both variants on HPCG-n"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi  # noqa: E402


def factors(ctx, name):
    if name.startswith("A"):
        lx, ly, lz = (int(v) for v in name[1:].split("x"))
        A = ctx.generate_anderson(lx, ly, lz)
    else:
        nx, ny, nz = (int(v) for v in name.split("x"))
        A = ctx.generate_hpcg(nx, ny, nz)
    n = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(n)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    return A, L, U, D, n


def check(ctx, name):
    A, L, U, D, n = factors(ctx, name)
    rng = np.random.default_rng(3)
    bh = rng.uniform(-1.0, 1.0, n)
    b = ctx.upload(bh)
    out = {}
    for variant in (5, 3):
        ctx.set_option("trsv_variant", variant)
        x, y, z, t = ctx.alloc(n), ctx.alloc(n), ctx.upload(bh), ctx.alloc(n)
        w0 = ctx.info()["wave_solves"]
        ctx.call("bis_sptrsv", L.h, x, D, b)
        ctx.call("bis_bsptrsv", U.h, y, D, b)
        ctx.call("bis_sptrsv", L.h, z, D, z)                      # in place
        ctx.call("bis_sptrsv", L.h, x, D, b)                      # again: the other working vector
        ctx.call("bis_apply_preconditioner", capi.PRECOND["sgs"], n, L.h, U.h, D, None, None, None, t, b, ctx.alloc(n), None)
        ctx.sync()
        out[variant] = [ctx.download(v, n) for v in (x, y, z, t)]
        used = ctx.info()["wave_solves"] - w0
        if variant == 5 and used == 0:
            print(f"{name}: variant 5 NOT used (not recognised as a stencil)")
            return False
    ok = all(np.array_equal(a, c) for a, c in zip(out[5], out[3]))
    if not ok:
        for what, a, c in zip(("forward", "backward", "in-place", "sgs-apply"), out[5], out[3]):
            bad = np.nonzero(a != c)[0]
            if bad.size:
                print(f"   {what}: {bad.size} entries differ, first at row {bad[0]} last at row {bad[-1]}; first few {bad[:6]}", flush=True)
    worst = max(float(np.max(np.abs(a - c))) for a, c in zip(out[5], out[3]))
    print(f"{name}: n={n} forward/backward/in-place/SGS-apply {'BIT-IDENTICAL' if ok else 'DIFFER'} (max abs diff {worst:.3e})", flush=True)
    ctx.set_option("trsv_variant", 0)
    for m in (L, U, A):
        m.free()
    return ok


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    with capi.Context(0) as ctx:
        if mode == "grids":          # python tools/run_trsv5.py grids 64x256x4 64x64x160 ...
            ok = True
            for nm in sys.argv[2:]:
                ok &= check(ctx, nm)
            print("TRSV5_GRIDS", "PASS" if ok else "FAIL")
            return 0 if ok else 1
        if mode == "check":
            names = ["8x8x8", "16x16x16", "20x14x11", "12x40x5", "33x70x3", "64x64x1", "7x9x1", "A12x10x8", "A40x36x7", "48x48x48"]
            ok = True
            for nm in names:
                ok &= check(ctx, nm)
            print("TRSV5_CHECK", "PASS" if ok else "FAIL")
            return 0 if ok else 1
        n = int(sys.argv[2])
        reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
        A, L, U, D, N = factors(ctx, f"{n}x{n}x{n}")
        b, x = ctx.upload(np.ones(N)), ctx.alloc(N)
        for variant in (5, 3):
            ctx.set_option("trsv_variant", variant)
            for T, fn in ((L, "bis_sptrsv"), (U, "bis_bsptrsv")):
                ctx.call(fn, T.h, x, D, b)
                ctx.sync()
                ctx.timer_start()
                for _ in range(reps):
                    ctx.call(fn, T.h, x, D, b)
                ms = ctx.timer_stop() / reps
                print(f"HPCG-{n} {fn} variant {variant}: {ms:.3f} ms per sweep ({1e3 * ms / (7 * n - 6):.3f} us per level of the wavefront)", flush=True)
        return 0


if __name__ == "__main__":
    sys.exit(main())
