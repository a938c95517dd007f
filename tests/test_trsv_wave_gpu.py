"""SpTRSV variant 5 (stencil wavefront, csrc/bis_sptrsv_wave.cuh; chosen automatically for 27-point factors, forced
here with trsv_variant = 5) must give the bits of the
dataflow solve (variant 3), which tests/test_kernels_gpu.py pins to the compiled reference: forward, backward,
in-place (x aliases b, gmres.hpp:288-291) and the SGS preconditioner with the D multiply folded into the forward
solve's store, on grids that exercise every edge of the scheme -- fewer lines than a warp, several 32-line blocks per
plane with a ragged last one, a single plane, a 7-point (Anderson) stencil, ILU(0) factors -- and the fallback."""
import numpy as np
import pytest

from basic_iterative_solvers_b200 import capi

pytestmark = pytest.mark.gpu

GRIDS = ["8x8x8", "20x14x11", "12x40x5", "33x70x3", "64x64x1", "7x9x1", "A12x10x8", "A40x36x7", "48x48x48",
         "64x256x4", "64x64x160", "32x161x6", "224x225x8"]


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


@pytest.mark.parametrize("name,tries", [("224x225x8", 40), ("256x256x20", 6)])
def test_wavefront_solve_repeated_at_eight_blocks_per_plane(ctx, name, tries):
    """Eight 32-line blocks per plane (16 warps per CTA) and long lines: the configuration in which a bulk copy of a
    later matrix record once overtook the loads of the current one (fixed by the proxy fence in refill()); the
    failure showed in ~1 % of the solves, at the last row of a line, so the solve is repeated."""
    A = _matrix(ctx, name)
    n = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(n)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    b, x = ctx.upload(np.random.default_rng(3).uniform(-1.0, 1.0, n)), ctx.alloc(n)
    want = {}
    try:
        ctx.set_option("trsv_variant", 3)
        for T, fn in ((L, "bis_sptrsv"), (U, "bis_bsptrsv")):
            ctx.call(fn, T.h, x, D, b)
            ctx.sync()
            want[fn] = ctx.download(x, n)
        ctx.set_option("trsv_variant", 5)
        for _ in range(tries):
            for T, fn in ((L, "bis_sptrsv"), (U, "bis_bsptrsv")):
                ctx.call(fn, T.h, x, D, b)
                ctx.sync()
                assert np.array_equal(ctx.download(x, n), want[fn]), fn
    finally:
        ctx.set_option("trsv_variant", 0)
    for m in (L, U, A):
        m.free()
    for v in (D, b, x):
        ctx.free(v)


def test_wavefront_is_chosen_automatically_for_27_point_factors_only(ctx):
    """trsv_variant = 0: the cost model picks the wavefront for HPCG factors (4 levels per plane) and the dataflow
    solve for 7-point factors (1 level per plane)."""
    for gen, expect in ((lambda: ctx.generate_hpcg(48, 48, 48), True), (lambda: ctx.generate_anderson(40, 36, 20), False)):
        A = gen()
        n = A.info()["n_rows"]
        L, U = ctx.split_triangular(A)
        D = ctx.alloc(n)
        ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
        b, x = ctx.upload(np.ones(n)), ctx.alloc(n)
        w0 = ctx.info()["wave_solves"]
        ctx.call("bis_sptrsv", L.h, x, D, b)
        ctx.call("bis_bsptrsv", U.h, x, D, b)
        ctx.sync()
        assert (ctx.info()["wave_solves"] - w0 == 2) == expect
        for m in (L, U, A):
            m.free()
        for v in (D, b, x):
            ctx.free(v)


@pytest.mark.parametrize("name", ["20x14x11", "33x70x3", "64x64x160", "224x225x8"])
def test_wavefront_cluster_sizes_give_the_same_bits(ctx, name):
    """Planes per thread-block cluster (option wave_cluster): 1 = every hand-over through L2, > 1 = planes of a cluster
    push their values into the next plane's shared memory.  Every size -- also those that do not divide the number of
    planes, and a change of size between two solves with the same factor (which planes leave a copy in the twin
    working vectors depends on it) -- must give the bits of the dataflow solve."""
    A = _matrix(ctx, name)
    n = A.info()["n_rows"]
    L, U = ctx.split_triangular(A)
    D = ctx.alloc(n)
    ctx.call("bis_matrix_extract_diagonal", A.h, D, None)
    bh = np.random.default_rng(11).uniform(-1.0, 1.0, n)
    try:
        ctx.set_option("trsv_variant", 3)
        want = _solves(ctx, L, U, D, bh, n)
        ctx.set_option("trsv_variant", 5)
        for cl in (8, 1, 2, 16, 4, 8):
            ctx.set_option("wave_cluster", cl)
            got = _solves(ctx, L, U, D, bh, n)
            for g, w in zip(got, want):
                assert np.array_equal(g, w), f"wave_cluster = {cl}"
    finally:
        ctx.set_option("trsv_variant", 0)
        ctx.set_option("wave_cluster", 8)
    for m in (L, U, A):
        m.free()
    ctx.free(D)
