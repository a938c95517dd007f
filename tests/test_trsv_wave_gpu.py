"""SpTRSV variant 5 (stencil wavefront, csrc/bis_sptrsv_wave.cuh; opt-in: trsv_variant = 5) must give the bits of the
dataflow solve (variant 3), which tests/test_kernels_gpu.py pins to the compiled reference: forward, backward,
in-place (x aliases b, gmres.hpp:288-291) and the SGS preconditioner with the D multiply folded into the forward
solve's store, on grids that exercise every edge of the scheme -- fewer lines than a warp, several 32-line blocks per
plane with a ragged last one, a single plane, a 7-point (Anderson) stencil, ILU(0) factors -- and the fallback."""
import numpy as np
import pytest

from basic_iterative_solvers_b200 import capi

pytestmark = pytest.mark.gpu

GRIDS = ["8x8x8", "20x14x11", "12x40x5", "33x70x3", "64x64x1", "7x9x1", "A12x10x8", "A40x36x7", "48x48x48"]


def _matrix(ctx, name):
    if name.startswith("A"):
        lx, ly, lz = (int(v) for v in name[1:].split("x"))
        return ctx.generate_anderson(lx, ly, lz)
    nx, ny, nz = (int(v) for v in name.split("x"))
    return ctx.generate_hpcg(nx, ny, nz)


def _solves(ctx, L, U, D, bh, n):
    b = ctx.upload(bh)
    x, y, z, t, tmp = ctx.alloc(n), ctx.alloc(n), ctx.upload(bh), ctx.alloc(n), ctx.alloc(n)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.call("bis_bsptrsv", U.h, y, D, b)
    ctx.call("bis_sptrsv", L.h, z, D, z)                      # in place
    ctx.call("bis_sptrsv", L.h, x, D, b)                      # again: the other working vector
    ctx.call("bis_apply_preconditioner", capi.PRECOND["sgs"], n, L.h, U.h, D, None, None, None, t, b, tmp, None)
    ctx.sync()
    out = [ctx.download(v, n) for v in (x, y, z, t)]
    for v in (b, x, y, z, t, tmp):
        ctx.free(v)
    return out


@pytest.mark.parametrize("name", GRIDS)
def test_wavefront_solve_is_bit_identical_to_dataflow_solve(ctx, name):
    A = _matrix(ctx, name)
    n = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(n)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    bh = np.random.default_rng(3).uniform(-1.0, 1.0, n)
    try:
        ctx.set_option("trsv_variant", 5)
        w0 = ctx.info()["wave_solves"]
        got = _solves(ctx, L, U, D, bh, n)
        assert ctx.info()["wave_solves"] - w0 == 6, "the stencil wavefront did not serve the solves"
        ctx.set_option("trsv_variant", 3)
        want = _solves(ctx, L, U, D, bh, n)
    finally:
        ctx.set_option("trsv_variant", 0)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    for m in (L, U, A):
        m.free()
    ctx.free(D)


def test_wavefront_solve_ilu0_factors(ctx):
    """ILU(0) factors of a 27-point matrix keep its pattern: served by the wavefront, same bits."""
    A = ctx.generate_hpcg(18, 37, 6)
    n = A.info()["n_rows"]
    L, U, LD, UD = ctx.ilu0(A, n)
    bh = np.random.default_rng(5).uniform(-1.0, 1.0, n)
    out = {}
    try:
        for variant in (5, 3):
            ctx.set_option("trsv_variant", variant)
            b, t, o = ctx.upload(bh), ctx.alloc(n), ctx.alloc(n)
            ctx.call("bis_sptrsv", L.h, t, LD, b)
            ctx.call("bis_bsptrsv", U.h, o, UD, t)
            ctx.sync()
            out[variant] = (ctx.download(t, n), ctx.download(o, n))
    finally:
        ctx.set_option("trsv_variant", 0)
    assert np.array_equal(out[5][0], out[3][0]) and np.array_equal(out[5][1], out[3][1])


def test_wavefront_rejects_unstructured_factor(ctx):
    """A factor that is no stencil cannot be forced onto variant 5: the call fails loudly."""
    rng = np.random.default_rng(9)
    n = 400
    rows, cols = [], []
    for r in range(1, n):
        for c in sorted(set(rng.integers(0, r, size=min(r, 3)).tolist())):
            rows.append(r)
            cols.append(c)
    rp = np.zeros(n + 1, np.int32)
    np.add.at(rp, np.array(rows) + 1, 1)
    rp = np.cumsum(rp).astype(np.int32)
    L = ctx.upload_triangular(rp, np.array(cols, np.int32), rng.uniform(-1, 1, len(cols)), upper=False)
    D, b, x = ctx.upload(np.full(n, 4.0)), ctx.upload(np.ones(n)), ctx.alloc(n)
    try:
        ctx.set_option("trsv_variant", 5)
        with pytest.raises(capi.BisError):
            ctx.call("bis_sptrsv", L.h, x, D, b)
    finally:
        ctx.set_option("trsv_variant", 0)
    ctx.call("bis_sptrsv", L.h, x, D, b)     # the dataflow solve serves it
    ctx.sync()
