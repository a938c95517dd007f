"""GPU parity tests, kernel level: every kernels.hpp entry of the C-ABI (include/bis_b200.h)
against the oracle (oracle/port) and the fixtures produced by the compiled reference.

Bit-exact: triangular solves, every preconditioner built from them, the axpby family, scale,
copy, init, normalize_x given the same SpMV input, ILU(0) applies, generators, diagonal peel.
Also bit-exact: SpMV in its default variant (TMA-staged tiles, one thread per row, products added
in storage order with separate roundings -- the order the compiled reference uses), and with it the
Jacobi sweep, compute_residual and b - T x.
Tolerance 1e-13 relative: dot and norm (the reference sums sequentially, a GPU cannot), and the
vector-CRS SpMV variant kept for rows too long for a shared-memory tile.
"""
import numpy as np
import pytest

from conftest import golden
from oracle import matgen, port

from basic_iterative_solvers_b200 import capi

pytestmark = pytest.mark.gpu

RED_TOL = 1e-13


def rel(a, b):
    s = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) / s


def i32(*a):
    return np.array(a, np.int32)


def f64(*a):
    return np.array(a, np.float64)


def run_spmv(ctx, rp, col, val, x, lanes=0, variant=None, **opts):
    """variant 0 (default): auto = 3 (windowed x, all operands by TMA) when the matrix is representable,
    else 2 (TMA-staged CRS tiles + x gathers); both add the products in storage order: bit-exact.
    lanes > 0: variant 1, vector CRS with that many lanes per row (tolerance only)."""
    ctx.set_option("spmv_variant", (1 if lanes else 0) if variant is None else variant)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.set_option("spmv_lanes", lanes)
    A = ctx.upload_crs(rp, col, val)      # the tile format (variant 3) is built at upload with these options
    dx, dy = ctx.upload(x), ctx.alloc(len(rp) - 1)
    ctx.call("bis_spmv", A.h, dx, dy)
    y = ctx.download(dy, len(rp) - 1)
    ctx.set_option("spmv_lanes", 0)
    ctx.set_option("spmv_variant", 0)
    for k in opts:
        ctx.set_option(k, 0)
    ctx.free(dx), ctx.free(dy), A.free()
    return y


# ---- reference KATs through the C-ABI (tests/test_kernels.cpp) -------------------------------
def test_kat_spmv(ctx):
    assert np.array_equal(run_spmv(ctx, i32(0, 1, 2, 3), i32(0, 1, 2), f64(2, 3, 4), f64(1, 2, 3)), f64(2, 6, 12))
    y = run_spmv(ctx, i32(0, 3, 6, 9), i32(0, 1, 2, 0, 1, 2, 0, 1, 2), f64(*range(1, 10)), f64(1, 2, 3))
    assert np.array_equal(y, f64(14, 32, 50))


def test_kat_sptrsv(ctx):
    L = ctx.upload_triangular(i32(0, 0, 1, 3), i32(0, 0, 1), f64(1, -2, 1), upper=False)
    U = ctx.upload_triangular(i32(0, 2, 3, 3), i32(1, 2, 2), f64(1, -2, 1), upper=True)
    D = ctx.upload(f64(2, 3, 4))
    x = ctx.alloc(3)
    b = ctx.upload(f64(2, 7, 12))
    ctx.call("bis_sptrsv", L.h, x, D, b)
    assert np.allclose(ctx.download(x, 3), f64(1, 2, 3), atol=1e-9)
    ctx.upload(f64(-2, 9, 12), b)
    ctx.call("bis_bsptrsv", U.h, x, D, b)
    assert np.allclose(ctx.download(x, 3), f64(1, 2, 3), atol=1e-9)
    # wrong factor kind is an error, not a silent wrong answer
    with pytest.raises(capi.BisError):
        ctx.call("bis_sptrsv", U.h, x, D, b)
    L.free(), U.free()


def test_kat_vector_ops(ctx):
    a, b, out = ctx.upload(f64(1, 2, 3)), ctx.upload(f64(4, 5, 6)), ctx.alloc(3)
    ctx.call("bis_sum_vectors", out, a, b, 3, 2.0)
    assert np.array_equal(ctx.download(out, 3), f64(9, 12, 15))
    ctx.call("bis_subtract_vectors", out, a, b, 3, 2.0)
    assert np.array_equal(ctx.download(out, 3), f64(-7, -8, -9))
    ctx.call("bis_elemwise_mult_vectors", out, a, b, 3, 1.0)
    assert np.array_equal(ctx.download(out, 3), f64(4, 10, 18))
    ctx.call("bis_elemwise_div_vectors", out, b, a, 3, 1.0)
    assert np.array_equal(ctx.download(out, 3), f64(4, 2.5, 2))
    c = ctx.upload(f64(2, 4, 5))
    assert ctx.dot(a, c, 3) == 25.0
    ctx.call("bis_scale", out, a, 3.0, 3)
    assert np.array_equal(ctx.download(out, 3), f64(3, 6, 9))
    d = ctx.upload(f64(3, 4))
    assert ctx.norm(d, 2) == 5.0
    assert ctx.norm(d, 0) == 0.0          # tests/test_utilities.cpp:55-62 (empty vector)
    ctx.call("bis_init_vector", out, 7.5, 3)
    assert np.array_equal(ctx.download(out, 3), f64(7.5, 7.5, 7.5))
    ctx.call("bis_copy_vector", out, a, 3)
    assert np.array_equal(ctx.download(out, 3), f64(1, 2, 3))


# ---- SpMV ------------------------------------------------------------------------------
@pytest.mark.parametrize("lanes", [0, 2, 4, 8, 16, 32])
def test_spmv_hpcg_vs_oracle(ctx, lanes):
    rp, col, val = matgen.hpcg(24, 20, 17)
    x = np.random.default_rng(1).uniform(-1, 1, len(rp) - 1)
    got, want = run_spmv(ctx, rp, col, val, x, lanes), port.spmv(rp, col, val, x)
    if lanes == 0:
        assert np.array_equal(got, want)       # in-order sums: the reference's bits
    else:
        assert rel(got, want) <= RED_TOL


@pytest.mark.parametrize("rows,stages,smem_kb", [(256, 2, 200), (128, 4, 200), (64, 3, 200), (256, 0, 170), (128, 2, 100), (64, 0, 50), (32, 0, 0)])
def test_spmv_tma_tile_shapes_bit_exact(ctx, rows, stages, smem_kb):
    """Every tile shape / pipeline depth of the TMA-staged variant gives the same bits, on a
    ragged matrix whose tiles start and end off the 16-byte copy alignment, with empty rows, an
    empty leading tile and a row count that is not a multiple of the tile."""
    rng = np.random.default_rng(12)
    n = 70001
    lens = rng.integers(0, 28, n)
    lens[:600] = 0
    lens[::11] = 0
    for rp_dtype in (np.int32, np.int64):
        rp = np.zeros(n + 1, rp_dtype)
        np.cumsum(lens, out=rp[1:])
        col = rng.integers(0, n, rp[-1]).astype(np.int32)
        val = rng.uniform(-1, 1, rp[-1])
        x = rng.uniform(-1, 1, n)
        got = run_spmv(ctx, rp, col, val, x, 0, variant=2, spmv_rows=rows, spmv_stages=stages, spmv_smem_kb=smem_kb)
        assert np.array_equal(got, port.spmv(rp.astype(np.int32), col, val, x))


@pytest.mark.parametrize("rows,stages,smem_kb", [(0, 0, 0), (32, 2, 0), (64, 3, 0), (128, 2, 0), (256, 2, 200), (128, 4, 200)])
@pytest.mark.parametrize("dims", [(24, 20, 17), (7, 5, 3), (130, 3, 2)])
def test_spmv_windowed_bit_exact(ctx, rows, stages, smem_kb, dims):
    """Variant 3 (x windows + 16-bit local column ids, producer warp + mbarrier full/empty pipeline) on
    stencil matrices, forced (spmv_variant=3 fails loudly if the matrix is not representable): every tile
    shape and pipeline depth gives the reference's bits, 32- and 64-bit row_ptr, odd sizes."""
    rp, col, val = matgen.hpcg(*dims)
    x = np.random.default_rng(3).uniform(-1, 1, len(rp) - 1)
    want = port.spmv(rp, col, val, x)
    for r in (rp, rp.astype(np.int64)):
        got = run_spmv(ctx, r, col, val, x, 0, variant=3, win_rows=rows, spmv_stages=stages, spmv_smem_kb=smem_kb)
        assert np.array_equal(got, want)


def test_spmv_windowed_anderson_periodic_and_fallback(ctx):
    """Periodic Anderson rows wrap around (far-away columns: more windows per tile); unstructured
    matrices are not representable: auto falls back to variant 2, forcing variant 3 is an error."""
    rp, col, val = matgen.anderson(12, 10, 9, 5.0, 1.0, 7, True)
    x = np.random.default_rng(5).uniform(-1, 1, len(rp) - 1)
    assert np.array_equal(run_spmv(ctx, rp, col, val, x, 0, variant=3), port.spmv(rp, col, val, x))
    rng = np.random.default_rng(6)
    n = 4000
    lens = rng.integers(1, 20, n)
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    col = rng.integers(0, n, rp[-1]).astype(np.int32)
    val = rng.uniform(-1, 1, rp[-1])
    x = rng.uniform(-1, 1, n)
    assert np.array_equal(run_spmv(ctx, rp, col, val, x, 0), port.spmv(rp, col, val, x))
    with pytest.raises(capi.BisError):
        run_spmv(ctx, rp, col, val, x, 0, variant=3)
    ctx.set_option("spmv_variant", 0)


def test_spmv_ragged_empty_rows_and_64bit_rowptr(ctx):
    rng = np.random.default_rng(2)
    n = 5000
    lens = rng.integers(0, 70, n)
    lens[::7] = 0
    lens[13] = 900                     # one very long row
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    col = rng.integers(0, n, rp[-1]).astype(np.int32)
    val = rng.uniform(-1, 1, rp[-1])
    x = rng.uniform(-1, 1, n)
    want = port.spmv(rp.astype(np.int32), col, val, x)
    for r in (rp, rp.astype(np.int32)):
        got = run_spmv(ctx, r, col, val, x)
        assert rel(got, want) <= RED_TOL
        assert np.all(got[lens == 0] == 0.0)
    # empty matrix
    A = ctx.upload_crs(i32(0), i32(), f64())
    assert A.info()["n_rows"] == 0
    A.free()


@pytest.mark.parametrize("name", ["fdm2d16", "band_klein", "hpcg16"])
def test_spmv_and_fused_forms_vs_reference_fixture(ctx, name):
    g = golden(name)
    rp, col, val = matgen.hpcg(16) if name == "hpcg16" else (g["rp"], g["col"], g["val"])
    x, v = g["k__x"], g["k__v"]
    n = x.size
    assert np.array_equal(run_spmv(ctx, rp, col, val, x), g["k__spmv"])     # the compiled reference's bits
    A = ctx.upload_crs(rp, col, val)
    dx, dv, dy, dr = ctx.upload(x), ctx.upload(v), ctx.alloc(n), ctx.alloc(n)
    y_ref = port.spmv(rp, col, val, x)
    # spmv + dots
    ctx.call("bis_spmv_dot", A.h, dx, dy, dv, 20, 21)
    s = ctx.scalars(20, 2)
    assert np.array_equal(ctx.download(dy, n), y_ref)
    assert abs(s[0] - y_ref @ v) <= 1e-12 * np.linalg.norm(y_ref) * np.linalg.norm(v)
    assert abs(s[1] - y_ref @ y_ref) <= 1e-12 * (y_ref @ y_ref)
    # residual + squared norm (compute_residual, kernels.hpp:155-162)
    ctx.call("bis_spmv_residual", A.h, dx, dv, dr, dy, 22)
    r_ref = v - y_ref
    assert np.array_equal(ctx.download(dr, n), r_ref)
    assert abs(ctx.scalars(22)[0] - r_ref @ r_ref) <= 1e-12 * (r_ref @ r_ref)
    ctx.call("bis_compute_residual", A.h, dx, dv, dr, dy)
    assert np.array_equal(ctx.download(dr, n), r_ref) and np.array_equal(ctx.download(dy, n), y_ref)
    # Jacobi sweep == spmv then normalize_x (jacobi.hpp:27-52); fused and unfused agree bit for bit
    f = port.factor(rp, col, val, "sgs")
    dD, dxn = ctx.upload(f.A_D), ctx.alloc(n)
    ctx.call("bis_spmv_jacobi", A.h, dD, dv, dx, dxn)
    fused = ctx.download(dxn, n)
    ctx.call("bis_spmv", A.h, dx, dxn)
    ctx.call("bis_normalize_x", dxn, dx, dD, dv, n)
    assert np.array_equal(fused, ctx.download(dxn, n))
    assert np.array_equal(fused, g["k__normalize_x"])
    # sweep + residual of the starting iterate from ONE product (what JacobiSolver::iterate enqueues): the sweep's
    # bits, the residual vector's bits, and the squared norm of the separate residual kernel bit for bit
    ctx.call("bis_spmv_residual", A.h, dx, dv, dr, None, 22)
    rr_sep = ctx.scalars(22)[0]
    dxn2, dr2 = ctx.alloc(n), ctx.alloc(n)
    ctx.call("bis_spmv_jacobi_residual", A.h, dD, dv, dx, dxn2, dr2, 23)
    assert np.array_equal(ctx.download(dxn2, n), fused)
    assert np.array_equal(ctx.download(dr2, n), r_ref)
    assert ctx.scalars(23)[0] == rr_sep
    # b - T x on a strictly triangular factor (gauss_seidel.hpp:30-34)
    U = ctx.upload_triangular(f.u_rp, f.u_col, f.u_val, upper=True)
    ctx.call("bis_spmv_sub", U.h, dx, dv, dr)
    assert np.array_equal(ctx.download(dr, n), v - port.spmv(f.u_rp, f.u_col, f.u_val, x))
    U.free(), A.free()


# ---- triangular solves and preconditioners: bit-exact ------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("name", ["fdm2d16", "band_klein", "hpcg16"])
def test_triangular_solves_bit_exact(ctx, name, variant):
    g = golden(name)
    rp, col, val = matgen.hpcg(16) if name == "hpcg16" else (g["rp"], g["col"], g["val"])
    x = g["k__x"]
    n = x.size
    f = port.factor(rp, col, val, "sgs")
    fi = port.factor(rp, col, val, "ilu0")
    ctx.set_option("trsv_variant", variant)
    try:
        L = ctx.upload_triangular(f.l_rp, f.l_col, f.l_val, upper=False)
        U = ctx.upload_triangular(f.u_rp, f.u_col, f.u_val, upper=True)
        dD, db, dx = ctx.upload(f.A_D), ctx.upload(x), ctx.alloc(n)
        ctx.call("bis_sptrsv", L.h, dx, dD, db)
        assert np.array_equal(ctx.download(dx, n), g["k__sptrsv"])
        ctx.call("bis_bsptrsv", U.h, dx, dD, db)
        assert np.array_equal(ctx.download(dx, n), g["k__bsptrsv"])
        # in place: x aliases b (gmres.hpp:288-291, bicgstab.hpp:157-160)
        ctx.call("bis_copy_vector", dx, db, n)
        ctx.call("bis_sptrsv", L.h, dx, dD, dx)
        assert np.array_equal(ctx.download(dx, n), g["k__sptrsv_inplace"])
        dAi, dLD, dUD = ctx.upload(f.A_D_inv), ctx.upload(np.ones(n)), ctx.upload(np.ones(n))
        dtmp, dwork, dout = ctx.alloc(n), ctx.alloc(n), ctx.alloc(n)
        for pre in ("none", "j", "gs", "bgs", "sgs"):
            ctx.call("bis_apply_preconditioner", capi.PRECOND[pre], n, L.h, U.h, dD, dAi, dLD, dUD,
                     dout, db, dtmp, dwork)
            assert np.array_equal(ctx.download(dout, n), g["k__precond__" + pre]), pre
        # in-place application (gmres.hpp:173-176)
        ctx.call("bis_copy_vector", dout, db, n)
        ctx.call("bis_apply_preconditioner", capi.PRECOND["sgs"], n, L.h, U.h, dD, dAi, dLD, dUD,
                 dout, dout, dtmp, dwork)
        assert np.array_equal(ctx.download(dout, n), g["k__precond__sgs"])
        # two-stage GS with PRECOND_INNER_ITERS = 0 (kernels.hpp:375-385) against the oracle
        for pre in ("2st", "s2st"):
            ctx.call("bis_apply_preconditioner", capi.PRECOND[pre], n, L.h, U.h, dD, dAi, dLD, dUD,
                     dout, db, dtmp, dwork)
            assert np.array_equal(ctx.download(dout, n), port.apply_preconditioner(pre, f, x)), pre
        L.free(), U.free()
        # ILU(0): L_D == 1, U_D from the factorisation
        L = ctx.upload_triangular(fi.l_rp, fi.l_col, fi.l_val, upper=False)
        U = ctx.upload_triangular(fi.u_rp, fi.u_col, fi.u_val, upper=True)
        ctx.upload(fi.L_D, dLD), ctx.upload(fi.U_D, dUD)
        ctx.call("bis_apply_preconditioner", capi.PRECOND["ilu0"], n, L.h, U.h, dD, dAi, dLD, dUD,
                 dout, db, dtmp, dwork)
        assert np.array_equal(ctx.download(dout, n), g["k__precond__ilu0"])
        L.free(), U.free()
    finally:
        ctx.set_option("trsv_variant", 0)
    ctx.sync()


def test_triangular_solve_long_rows_and_many_levels(ctx):
    """Rows longer than the register prefetch window, a pure chain (n levels) and a diagonal
    matrix (one level): the schedule never changes the bits."""
    rng = np.random.default_rng(4)
    n = 3000
    cols, rp = [], [0]
    for r in range(n):
        k = min(r, int(rng.integers(0, 40)))
        c = np.sort(rng.choice(r, k, replace=False)) if k else np.zeros(0, int)
        if r > 0 and r % 3 == 0 and (r - 1) not in c:
            c = np.sort(np.append(c, r - 1))       # long dependency chains
        cols.append(c)
        rp.append(rp[-1] + len(c))
    col = np.concatenate(cols).astype(np.int32)
    rp = np.array(rp, np.int32)
    val = rng.uniform(-0.05, 0.05, col.size)
    D = rng.uniform(1, 2, n)
    b = rng.uniform(-1, 1, n)
    L = ctx.upload_triangular(rp, col, val, upper=False)
    assert L.info()["n_levels"] > 100
    dD, db, dx = ctx.upload(D), ctx.upload(b), ctx.alloc(n)
    ctx.call("bis_sptrsv", L.h, dx, dD, db)
    assert np.array_equal(ctx.download(dx, n), port.sptrsv(rp, col, val, D, b))
    L.free()
    # mirrored into an upper factor
    dense_rows = np.repeat(np.arange(n), np.diff(rp))
    order = np.lexsort((dense_rows, col))          # transpose: sort by (col, row)
    urp = np.zeros(n + 1, np.int32)
    np.add.at(urp, col + 1, 1)
    urp = np.cumsum(urp).astype(np.int32)
    ucol, uval = dense_rows[order].astype(np.int32), val[order]
    U = ctx.upload_triangular(urp, ucol, uval, upper=True)
    ctx.call("bis_bsptrsv", U.h, dx, dD, db)
    assert np.array_equal(ctx.download(dx, n), port.sptrsv(urp, ucol, uval, D, b, backward=True))
    U.free()
    # chain: n levels of one row
    m = 2000
    crp = np.arange(0, m, dtype=np.int32)
    crp = np.concatenate([[0], crp]).astype(np.int32)
    ccol = np.arange(0, m - 1, dtype=np.int32)
    cval = np.full(m - 1, -0.5)
    C = ctx.upload_triangular(crp, ccol, cval, upper=False)
    assert C.info()["n_levels"] == m
    dD2, db2, dx2 = ctx.upload(np.full(m, 1.5)), ctx.upload(np.ones(m)), ctx.alloc(m)
    ctx.call("bis_sptrsv", C.h, dx2, dD2, db2)
    assert np.array_equal(ctx.download(dx2, m), port.sptrsv(crp, ccol, cval, np.full(m, 1.5), np.ones(m)))
    C.free()
    # diagonal only: empty strict factor
    E = ctx.upload_triangular(np.zeros(m + 1, np.int32), i32(), f64(), upper=False)
    ctx.call("bis_sptrsv", E.h, dx2, dD2, db2)
    assert np.array_equal(ctx.download(dx2, m), np.ones(m) / 1.5)
    E.free()
    # entries on the wrong side of the diagonal are rejected at upload
    with pytest.raises(capi.BisError):
        ctx.upload_triangular(i32(0, 1, 1), i32(1), f64(1.0), upper=False)
    ctx.sync()


# ---- BLAS-1 and fused vector kernels -----------------------------------------------------
def test_axpby_family_bit_exact_with_aliasing(ctx):
    g = golden("hpcg16")
    x, v = g["k__x"], g["k__v"]
    n = x.size
    dx, dv, out = ctx.upload(x), ctx.upload(v + 2.0), ctx.alloc(n)
    for fn in ("subtract_vectors", "sum_vectors", "elemwise_mult_vectors", "elemwise_div_vectors"):
        ctx.call("bis_" + fn, out, dx, dv, n, 0.37)
        assert np.array_equal(ctx.download(out, n), g["k__" + fn]), fn
        # out aliases a, then out aliases b (gauss_seidel.hpp:34, gmres.hpp:25, kernels.hpp:369)
        ctx.call("bis_copy_vector", out, dx, n)
        ctx.call("bis_" + fn, out, out, dv, n, 0.37)
        assert np.array_equal(ctx.download(out, n), g["k__" + fn]), fn + " out=a"
        ctx.call("bis_copy_vector", out, dv, n)
        ctx.call("bis_" + fn, out, dx, out, n, 0.37)
        assert np.array_equal(ctx.download(out, n), g["k__" + fn]), fn + " out=b"
    ctx.call("bis_scale", out, dx, -1.25, n)
    assert np.array_equal(ctx.download(out, n), g["k__scale"])
    dvv = ctx.upload(v)
    assert abs(ctx.dot(dx, dvv, n) - g["k__dot"][0]) <= RED_TOL * np.sqrt(n)
    assert abs(ctx.norm(dx, n) - g["k__norm"][0]) <= RED_TOL * g["k__norm"][0]


@pytest.mark.parametrize("n", [1, 31, 1000, 4097, 1 << 20, (1 << 22) + 3])
def test_reductions_deterministic_and_accurate(ctx, n):
    rng = np.random.default_rng(n)
    a, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    da, db = ctx.upload(a), ctx.upload(b)
    d1, d2 = ctx.dot(da, db, n), ctx.dot(da, db, n)
    assert d1 == d2                                   # fixed launch shape => bit-reproducible
    exact = float(np.sum(a.astype(np.longdouble) * b.astype(np.longdouble)))
    assert abs(d1 - exact) <= 1e-13 * np.sqrt(n) + 1e-15 * n ** 0.5
    nv = ctx.norm(da, n)
    assert abs(nv - np.linalg.norm(a)) <= 1e-13 * nv
    ctx.free(da), ctx.free(db)


def test_cg_fused_kernels_match_unfused_sequence(ctx):
    """bis_cg_update / bis_cg_direction against the reference's sequence of separate kernels
    (cg.hpp:19-52) run through the same C-ABI: identical per-element arithmetic => identical bits
    for the vectors; the fused reductions agree with separate dots to rounding."""
    rng = np.random.default_rng(9)
    n = 100003
    h = {k: rng.uniform(-1, 1, n) for k in ("x", "p", "r", "Ap")}
    D = rng.uniform(1, 2, n)
    d = {k: ctx.upload(v) for k, v in h.items()}
    dD = ctx.upload(D)
    xn, rn, zn, pn, t = (ctx.alloc(n) for _ in range(5))
    rz, pAp = 0.731, 1.93
    alpha = rz / pAp
    for pre in ("none", "j", "sgs"):
        ctx.call("bis_scalar_set", 2, rz)
        ctx.call("bis_scalar_set", 3, pAp)
        ctx.call("bis_cg_update", capi.PRECOND[pre], n, xn, d["x"], d["p"], rn, d["r"], d["Ap"], zn, dD,
                 2, 3, 0, 1)
        got_x, got_r = ctx.download(xn, n), ctx.download(rn, n)
        ctx.call("bis_sum_vectors", t, d["x"], d["p"], n, alpha)
        assert np.array_equal(got_x, ctx.download(t, n))
        ctx.call("bis_subtract_vectors", t, d["r"], d["Ap"], n, alpha)
        want_r = ctx.download(t, n)
        assert np.array_equal(got_r, want_r)
        s = ctx.scalars(0, 2)
        assert abs(s[0] - want_r @ want_r) <= 1e-12 * (want_r @ want_r)
        if pre == "none":
            assert np.array_equal(ctx.download(zn, n), want_r) and s[1] == s[0]
        if pre == "j":
            ctx.call("bis_elemwise_div_vectors", t, rn, dD, n, 1.0)
            want_z = ctx.download(t, n)
            assert np.array_equal(ctx.download(zn, n), want_z)
            assert abs(s[1] - want_r @ want_z) <= 1e-12 * abs(want_r @ want_z)
    ctx.call("bis_scalar_set", 1, 0.9)
    ctx.call("bis_scalar_set", 2, 0.4)
    ctx.call("bis_cg_direction", n, pn, zn, d["p"], 1, 2)
    ctx.call("bis_sum_vectors", t, zn, d["p"], n, 0.9 / 0.4)
    assert np.array_equal(ctx.download(pn, n), ctx.download(t, n))
    # the pair the host stack uses: update without x, then direction + x in one pass over p_old
    ctx.call("bis_scalar_set", 2, rz)
    ctx.call("bis_scalar_set", 3, pAp)
    rn2, zn2, xn2, pn2 = (ctx.alloc(n) for _ in range(4))
    ctx.call("bis_cg_update", capi.PRECOND["j"], n, None, None, None, rn2, d["r"], d["Ap"], zn2, dD, 2, 3, 0, 1)
    ctx.call("bis_cg_update", capi.PRECOND["j"], n, xn, d["x"], d["p"], rn, d["r"], d["Ap"], zn, dD, 2, 3, 4, 5)
    assert np.array_equal(ctx.download(rn2, n), ctx.download(rn, n)) and np.array_equal(ctx.download(zn2, n), ctx.download(zn, n))
    assert np.array_equal(ctx.scalars(0, 2), ctx.scalars(4, 2))
    ctx.call("bis_cg_direction_x", n, pn2, zn2, d["p"], xn2, d["x"], 1, 2, 3)
    beta = ctx.scalars(1, 1)[0] / rz
    ctx.call("bis_sum_vectors", t, zn2, d["p"], n, beta)
    assert np.array_equal(ctx.download(pn2, n), ctx.download(t, n))
    assert np.array_equal(ctx.download(xn2, n), ctx.download(xn, n))


def test_bicgstab_and_gmres_fused_kernels(ctx):
    rng = np.random.default_rng(10)
    n = 50021
    names = ("x", "y", "st", "s", "z", "r0", "r", "v", "p")
    h = {k: rng.uniform(-1, 1, n) for k in names}
    d = {k: ctx.upload(v) for k, v in h.items()}
    D = rng.uniform(1, 2, n)
    dD = ctx.upload(D)
    o1, o2, o3, t = (ctx.alloc(n) for _ in range(4))
    rho_old, r0v, zs, zz = 0.83, 1.7, 0.61, 2.3
    for slot, v in ((9, rho_old), (4, r0v), (5, zs), (6, zz)):
        ctx.call("bis_scalar_set", slot, v)
    alpha, omega = rho_old / r0v, zs / zz
    # s = r - alpha v ; s_tmp = s / D
    ctx.call("bis_bicgstab_s", capi.PRECOND["j"], n, o1, o2, d["r"], d["v"], dD, 9, 4)
    ctx.call("bis_subtract_vectors", t, d["r"], d["v"], n, alpha)
    s_want = ctx.download(t, n)
    assert np.array_equal(ctx.download(o1, n), s_want)
    ctx.call("bis_elemwise_div_vectors", t, o1, dD, n, 1.0)
    assert np.array_equal(ctx.download(o2, n), ctx.download(t, n))
    # x_new = (x + alpha y) + omega s_tmp ; r_new = s - omega z ; (r0,r_new) ; (r_new,r_new)
    ctx.call("bis_bicgstab_xr", n, o3, o1, d["x"], d["y"], d["st"], o2, d["s"], d["z"], d["r0"],
             9, 4, 5, 6, 7, 8)
    ctx.call("bis_sum_vectors", t, d["x"], d["y"], n, alpha)
    assert np.array_equal(ctx.download(o3, n), ctx.download(t, n))
    ctx.call("bis_sum_vectors", t, t, d["st"], n, omega)
    assert np.array_equal(ctx.download(o1, n), ctx.download(t, n))
    ctx.call("bis_subtract_vectors", t, d["s"], d["z"], n, omega)
    r_new = ctx.download(t, n)
    assert np.array_equal(ctx.download(o2, n), r_new)
    s = ctx.scalars(7, 2)
    assert abs(s[0] - h["r0"] @ r_new) <= 1e-12 * np.linalg.norm(r_new) * np.linalg.norm(h["r0"])
    assert abs(s[1] - r_new @ r_new) <= 1e-12 * (r_new @ r_new)
    # p_new = r_new + beta (p - omega v) ; y_next = p_new / D
    rho_new = s[0]
    beta = (rho_new / rho_old) * (alpha / omega)
    ctx.call("bis_bicgstab_p", capi.PRECOND["j"], n, o3, o1, d["p"], d["v"], o2, t, dD, 7, 9, 4, 5, 6)
    y_next = ctx.download(t, n)
    ctx.call("bis_subtract_vectors", t, d["p"], d["v"], n, omega)
    assert np.array_equal(ctx.download(o3, n), ctx.download(t, n))
    ctx.call("bis_sum_vectors", t, o2, t, n, beta)
    assert np.array_equal(ctx.download(o1, n), ctx.download(t, n))
    ctx.call("bis_elemwise_div_vectors", t, o1, dD, n, 1.0)
    assert np.array_equal(y_next, ctx.download(t, n))
    # MGS step: w -= h v_j ; (w, v_next)
    ctx.call("bis_scalar_set", 16, 0.37)
    ctx.call("bis_copy_vector", o1, d["x"], n)
    ctx.call("bis_mgs_step", n, o1, d["y"], d["z"], 16, 17)
    ctx.call("bis_subtract_vectors", t, d["x"], d["y"], n, 0.37)
    w = ctx.download(t, n)
    assert np.array_equal(ctx.download(o1, n), w)
    assert abs(ctx.scalars(17)[0] - w @ h["z"]) <= 1e-12 * np.linalg.norm(w) * np.linalg.norm(h["z"])
    ctx.call("bis_mgs_step", n, o1, d["y"], None, 16, 11)
    ctx.call("bis_subtract_vectors", t, t, d["y"], n, 0.37)
    w = ctx.download(t, n)
    assert abs(ctx.scalars(11)[0] - w @ w) <= 1e-12 * (w @ w)
    # basis normalisation: out = w * (1 / sqrt(sumsq))  (gmres.hpp:36-45)
    ctx.call("bis_scale_inv_norm", n, o2, o1, 11)
    ctx.call("bis_scale", t, o1, 1.0 / np.sqrt(ctx.scalars(11)[0]), n)
    assert np.array_equal(ctx.download(o2, n), ctx.download(t, n))
    # get_explicit_x: x = x_old + sum_j V_j y_j, left to right (gmres.hpp:326-375)
    k = 5
    V = rng.uniform(-1, 1, (k + 1, n))
    yv = rng.uniform(-1, 1, k)
    dV = ctx.upload(V.ravel())
    ctx.call("bis_gmres_update_x", n, k, dV, yv.ctypes.data, o1, d["x"], o2)
    acc = np.zeros(n)
    for j in range(k):
        acc = acc + V[j] * yv[j]
    assert np.array_equal(ctx.download(o2, n), acc)
    assert np.array_equal(ctx.download(o1, n), h["x"] + acc)


# ---- matrices: generators, diagonal, split -------------------------------------------------
def test_device_generators_match_numpy(ctx):
    for dims in ((16, 16, 16), (9, 7, 5), (1, 1, 1), (2, 1, 3), (33, 8, 4)):
        A = ctx.generate_hpcg(*dims)
        rp, col, val = A.download()
        w = matgen.hpcg(*dims, index_dtype=np.int64)
        assert np.array_equal(rp, w[0]) and np.array_equal(col, w[1]) and np.array_equal(val, w[2]), dims
        assert A.info()["nnz"] == matgen.hpcg_nnz(*dims)
        A.free()
    for periodic in (False, True):
        for dims in ((10, 8, 6), (3, 2, 1), (2, 2, 2)):
            A = ctx.generate_anderson(*dims, ranpot=5.0, t=1.0, seed=42, periodic=periodic)
            rp, col, val = A.download()
            w = matgen.anderson(*dims, 5.0, 1.0, 42, periodic)
            assert np.array_equal(rp, w[0]) and np.array_equal(col, w[1]) and np.array_equal(val, w[2])
            A.free()


def test_extract_diagonal_and_split(ctx):
    rp, col, val = matgen.anderson(9, 8, 7, 5.0, 1.0, 3, True)
    n = len(rp) - 1
    A = ctx.upload_crs(rp.astype(np.int32), col, val)
    f = port.factor(rp.astype(np.int32), col, val, "sgs")
    dD, dDi = ctx.alloc(n), ctx.alloc(n)
    ctx.call("bis_matrix_extract_diagonal", A.h, dD, dDi)
    assert np.array_equal(ctx.download(dD, n), f.A_D)
    assert np.array_equal(ctx.download(dDi, n), f.A_D_inv)
    L, U = ctx.split_triangular(A)
    lrp, lcol, lval = L.download()
    urp, ucol, uval = U.download()
    assert np.array_equal(lrp, f.l_rp) and np.array_equal(lcol, f.l_col) and np.array_equal(lval, f.l_val)
    assert np.array_equal(urp, f.u_rp) and np.array_equal(ucol, f.u_col) and np.array_equal(uval, f.u_val)
    L.free(), U.free(), A.free()
    # a missing diagonal is fatal in the reference (common.hpp:393-396)
    B = ctx.upload_crs(i32(0, 1, 2), i32(1, 0), f64(1, 1))
    with pytest.raises(capi.BisError):
        ctx.call("bis_matrix_extract_diagonal", B.h, dD, None)
    B.free()


def _host_levels(rp, col, upper):
    n = len(rp) - 1
    lev = np.zeros(n, np.int64)
    for r in (range(n - 1, -1, -1) if upper else range(n)):
        c = col[rp[r]:rp[r + 1]]
        lev[r] = (lev[c].max() + 1) if c.size else 0
    return lev


def _random_general(n, seed, zero_fraction=0.05):
    """Unsorted columns, a diagonal in every row, some stored zeros and some tiny diagonals: the cases the
    ILU(0) rules treat specially (LU_factors.hpp:370, 384, 410-412)."""
    rng = np.random.default_rng(seed)
    rows = []
    for r in range(n):
        k = int(rng.integers(1, 9))
        c = set(int(v) for v in rng.integers(max(0, r - 12), min(n, r + 13), size=k))
        c.add(r)
        c = np.array(sorted(c))
        rng.shuffle(c)
        v = rng.uniform(-1.0, 1.0, size=c.size)
        v[rng.uniform(size=c.size) < zero_fraction] = 0.0
        d = np.where(c == r)[0][0]
        v[d] = 4.0 + rng.uniform()
        if r % 37 == 5:
            v[d] = 1e-12      # below ILU0_PIVOT_TOLERANCE: replaced
        if r % 53 == 7:
            v[d] = -3e-9
        rows.append((c, v))
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = np.cumsum([len(c) for c, _ in rows])
    return rp, np.concatenate([c for c, _ in rows]).astype(np.int32), np.concatenate([v for _, v in rows])


@pytest.mark.parametrize("case", ["hpcg", "hpcg_rp64", "anderson", "random", "random_big"])
def test_device_ilu0_matches_host_factorisation(ctx, case):
    """bis_matrix_ilu0 (one dataflow launch) against the sequential host routine that tests/test_host_cpu.py
    pins to the compiled reference's factor_ILU0_old: every factor entry bit for bit."""
    if case in ("hpcg", "hpcg_rp64"):
        rp, col, val = matgen.hpcg(9, 7, 6)
    elif case == "anderson":
        rp, col, val = matgen.anderson(9, 8, 7, 5.0, 1.0, 3, True)
    else:
        rp, col, val = _random_general(300 if case == "random" else 20000, 11)
    rp = rp.astype(np.int32)
    n = len(rp) - 1
    f = port.factor(rp, col, val, "ilu0")
    # 64-bit row_ptr: the layout HPCG-512 needs on one GPU (bis_matrix_upload_crs64)
    A = ctx.upload_crs(rp.astype(np.int64) if case.endswith("rp64") else rp, col, val)
    L, U, dLD, dUD = ctx.ilu0(A, n)
    lrp, lcol, lval = L.download()
    urp, ucol, uval = U.download()
    assert np.array_equal(lrp, f.l_rp) and np.array_equal(lcol, f.l_col)
    assert np.array_equal(urp, f.u_rp) and np.array_equal(ucol, f.u_col)
    assert np.array_equal(lval.view(np.int64), f.l_val.view(np.int64))
    assert np.array_equal(uval.view(np.int64), f.u_val.view(np.int64))
    assert np.array_equal(ctx.download(dUD, n).view(np.int64), f.U_D.view(np.int64))
    assert np.array_equal(ctx.download(dLD, n), f.L_D)
    assert L.info()["n_levels"] == int(_host_levels(f.l_rp, f.l_col, False).max()) + 1
    assert U.info()["n_levels"] == int(_host_levels(f.u_rp, f.u_col, True).max()) + 1
    # and the factors solve: apply the preconditioner through the device level sets
    y = np.linspace(-1.0, 1.0, n)
    want = port.apply_preconditioner("ilu0", f, y)
    dy, dout, dtmp = ctx.upload(y), ctx.alloc(n), ctx.alloc(n)
    dAD = ctx.upload(f.A_D)
    ctx.call("bis_apply_preconditioner", capi.PRECOND["ilu0"], n, L.h, U.h, dAD, None, dLD, dUD, dout, dy, dtmp, None)
    assert np.array_equal(ctx.download(dout, n).view(np.int64), want.view(np.int64))
    for v in (dy, dout, dtmp, dAD, dLD, dUD):
        ctx.free(v)
    L.free(), U.free(), A.free()


def test_device_level_analysis_matches_host(ctx):
    rp, col, val = _random_general(5000, 5, zero_fraction=0.0)
    f = port.factor(rp, col, val, "sgs")
    A = ctx.upload_crs(rp, col, val)
    L, U = ctx.split_triangular(A)
    assert L.info()["n_levels"] == int(_host_levels(f.l_rp, f.l_col, False).max()) + 1
    assert U.info()["n_levels"] == int(_host_levels(f.u_rp, f.u_col, True).max()) + 1
    lrp, lcol, lval = L.download()
    assert np.array_equal(lrp, f.l_rp) and np.array_equal(lcol, f.l_col) and np.array_equal(lval, f.l_val)
    # both sweeps through the device-built level sets, bit for bit against the sequential oracle
    n = len(rp) - 1
    b = np.cos(np.arange(n))
    dD, db, dx = ctx.upload(f.A_D), ctx.upload(b), ctx.alloc(n)
    ctx.call("bis_sptrsv", L.h, dx, dD, db)
    assert np.array_equal(ctx.download(dx, n).view(np.int64), port.sptrsv(f.l_rp, f.l_col, f.l_val, f.A_D, b).view(np.int64))
    ctx.call("bis_bsptrsv", U.h, dx, dD, db)
    assert np.array_equal(ctx.download(dx, n).view(np.int64), port.sptrsv(f.u_rp, f.u_col, f.u_val, f.A_D, b, backward=True).view(np.int64))
    for v in (dD, db, dx):
        ctx.free(v)
    L.free(), U.free(), A.free()


def test_scale_symmetric_matches_numpy_restatement(ctx):
    # extract_scale + scale_mat (LU_factors.hpp:880-898, preprocessing.hpp:15-24): bit for bit
    rp, col, val = _random_general(4000, 3)
    A = ctx.upload_crs(rp, col, val)
    n = len(rp) - 1
    ds = ctx.alloc(n)
    ctx.call("bis_matrix_scale_symmetric", A.h, ds)
    want_val, want_s = port.scale_symmetric(rp, col, val)
    assert np.array_equal(ctx.download(ds, n).view(np.int64), want_s.view(np.int64))
    got = A.download()
    assert np.array_equal(got[2].view(np.int64), want_val.view(np.int64))
    # SpMV on the scaled matrix (the tile format does not hold values, so it stays valid)
    x = np.sin(np.arange(n))
    dx, dy = ctx.upload(x), ctx.alloc(n)
    ctx.call("bis_spmv", A.h, dx, dy)
    assert np.array_equal(ctx.download(dy, n).view(np.int64), port.spmv(rp, col, want_val, x).view(np.int64))
    for v in (ds, dx, dy):
        ctx.free(v)
    A.free()


@pytest.mark.parametrize("case", ["hpcg_64_64_8", "anderson_100_20_10", "hpcg_ilu0"])
def test_chain_variant_of_the_triangular_solve(ctx, case):
    """trsv_variant = 4 (bis_sptrsv_chain.cuh: a lane per chain of consecutively dependent rows, 32 chains per
    warp skewed by their levels): same bits as the sequential oracle, forward, backward, in place, and as
    SGS / ILU(0) preconditioner.  The grids are large enough for the chain format to be accepted."""
    if case == "anderson_100_20_10":
        rp, col, val = matgen.anderson(100, 20, 10, 5.0, 1.0, 3, False)
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        val = val.copy()
        val[rows == col] += 8.0
    else:
        rp, col, val = matgen.hpcg(64, 64, 8)
    rp = rp.astype(np.int32)
    n = len(rp) - 1
    pre = "ilu0" if case == "hpcg_ilu0" else "sgs"
    f = port.factor(rp, col, val, pre)
    DL, DU = (f.L_D, f.U_D) if pre == "ilu0" else (f.A_D, f.A_D)
    b = np.cos(np.arange(n) * 0.01) + 0.3
    ctx.set_option("trsv_variant", 4)
    try:
        before = ctx.info()["chain_solves"]
        L = ctx.upload_triangular(f.l_rp, f.l_col, f.l_val, upper=False)
        U = ctx.upload_triangular(f.u_rp, f.u_col, f.u_val, upper=True)
        dDL, dDU, db, dx = ctx.upload(DL), ctx.upload(DU), ctx.upload(b), ctx.alloc(n)
        for _ in range(2):      # the second solve reuses the format and the working vector
            ctx.call("bis_sptrsv", L.h, dx, dDL, db)
            assert np.array_equal(ctx.download(dx, n).view(np.int64),
                                  port.sptrsv(f.l_rp, f.l_col, f.l_val, DL, b).view(np.int64))
            ctx.call("bis_bsptrsv", U.h, dx, dDU, db)
            assert np.array_equal(ctx.download(dx, n).view(np.int64),
                                  port.sptrsv(f.u_rp, f.u_col, f.u_val, DU, b, backward=True).view(np.int64))
        # in place (x aliases b): gmres.hpp:288-291, bicgstab.hpp:157-160
        dxb = ctx.upload(b)
        ctx.call("bis_sptrsv", L.h, dxb, dDL, dxb)
        assert np.array_equal(ctx.download(dxb, n).view(np.int64), port.sptrsv(f.l_rp, f.l_col, f.l_val, DL, b).view(np.int64))
        # as a preconditioner
        want = port.apply_preconditioner(pre, f, b)
        dAD, dLD, dUD = ctx.upload(f.A_D), ctx.upload(f.L_D), ctx.upload(f.U_D)
        dout, dtmp = ctx.alloc(n), ctx.alloc(n)
        ctx.call("bis_apply_preconditioner", capi.PRECOND[pre], n, L.h, U.h, dAD, None, dLD, dUD, dout, db, dtmp, None)
        assert np.array_equal(ctx.download(dout, n).view(np.int64), want.view(np.int64))
        assert ctx.info()["chain_solves"] - before == 7, "the chain format was not accepted for this matrix"
        for v in (dDL, dDU, db, dx, dxb, dAD, dLD, dUD, dout, dtmp):
            ctx.free(v)
        L.free(), U.free()
    finally:
        ctx.set_option("trsv_variant", 0)


def test_zero_diagonal_is_fatal(ctx):
    # SanityChecker::zero_diag (common.hpp:388-391, LU_factors.hpp:842-845)
    B = ctx.upload_crs(i32(0, 2, 4), i32(0, 1, 0, 1), f64(1e-17, 1, 1, 2))
    d = ctx.alloc(2)
    with pytest.raises(capi.BisError, match="Zero detected on diagonal at row index 0"):
        ctx.call("bis_matrix_extract_diagonal", B.h, d, None)
    ctx.free(d)
    B.free()


def test_hpcg_levels_follow_wavefront(ctx):
    # HPCG natural ordering: level(x,y,z) = x + 2y + 4z  =>  7n - 6 levels
    n = 12
    A = ctx.generate_hpcg(n)
    L, U = ctx.split_triangular(A)
    assert L.info()["n_levels"] == 7 * n - 6 and U.info()["n_levels"] == 7 * n - 6
    L.free(), U.free(), A.free()


def test_bad_arguments_fail_loudly(ctx):
    with pytest.raises(capi.BisError):
        ctx.call("bis_dot_to_slot", None, None, 4, 4096)
    with pytest.raises(capi.BisError):
        ctx.upload_crs(i32(0, 2, 1), i32(0, 0), f64(1, 1))     # row_ptr not monotone / != nnz
    with pytest.raises(capi.BisError):
        ctx.set_option("no_such_option", 1)


def test_device_coo_to_crs_matches_reference_conversion(ctx):
    """bis_matrix_upload_coo = the reader's stable sort by row + convert_coo_to_crs (utilities.hpp:326-367): entries in
    file order (shuffled here, several per row, a symmetric pair, an empty row) must come out grouped by row with
    their order of appearance kept inside a row -- the summation order -- and SpMV on the result must give the
    oracle's bits on the same CRS."""
    rng = np.random.default_rng(41)
    n = 300
    lens = rng.integers(0, 9, size=n)
    lens[17] = 0
    I = np.repeat(np.arange(n), lens).astype(np.int32)
    J = rng.integers(0, n, size=I.size).astype(np.int32)
    V = rng.uniform(-1.0, 1.0, size=I.size)
    perm = rng.permutation(I.size)
    I, J, V = I[perm], J[perm], V[perm]
    order = np.argsort(I, kind="stable")
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, I + 1, 1)
    rp = np.cumsum(rp)
    A = ctx.upload_coo(n, n, I, J, V)
    got_rp, got_col, got_val = A.download()
    assert np.array_equal(got_rp, rp)
    assert np.array_equal(got_col, J[order]) and np.array_equal(got_val, V[order])
    x = rng.uniform(-1.0, 1.0, n)
    dx, dy = ctx.upload(x), ctx.alloc(n)
    ctx.call("bis_spmv", A.h, dx, dy)
    assert np.array_equal(ctx.download(dy, n), port.spmv(rp.astype(np.int32), J[order], V[order], x))
    # already grouped by row: no sort, same result
    B = ctx.upload_coo(n, n, I[order], J[order], V[order], sorted_by_row=True)
    b_rp, b_col, b_val = B.download()
    assert np.array_equal(b_rp, rp) and np.array_equal(b_col, J[order]) and np.array_equal(b_val, V[order])
    # a row index out of range is an error, as is an unsorted list passed as sorted
    with pytest.raises(capi.BisError):
        ctx.upload_coo(n, n, np.array([0, n], np.int32), np.array([0, 0], np.int32), np.array([1.0, 1.0]))
    with pytest.raises(capi.BisError):
        ctx.upload_coo(n, n, np.array([5, 2], np.int32), np.array([0, 0], np.int32), np.array([1.0, 1.0]), sorted_by_row=True)
    A.free()
    B.free()
