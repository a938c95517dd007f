"""Value dictionary of the windowed SpMV (csrc/bis_spmv.cu: win_build_dict): a matrix with at most 256 distinct
values streams a 1-byte index per nonzero instead of the 8-byte value and looks the value up in shared memory.  It is
lossless: y must have the bits of the SpMV that streams the values (and of native_spmv, kernels.hpp:22-42, which
tests/test_kernels_gpu.py pins), the matrices beyond 256 values must fall back, and values changed in place
(bis_matrix_scale_symmetric) must be picked up."""
import numpy as np
import pytest

from oracle import matgen, port

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _dictionary_on(ctx):
    """The dictionary is an option (spmv_vdict, off by default); a matrix gets one at its first SpMV while it is on."""
    ctx.set_option("spmv_vdict", 1)
    yield
    ctx.set_option("spmv_vdict", 0)


def _spmv(ctx, A, x):
    dx, dy = ctx.upload(x), ctx.alloc(x.size)
    ctx.call("bis_spmv", A.h, dx, dy)
    ctx.sync()
    y = ctx.download(dy, x.size)
    ctx.free(dx)
    ctx.free(dy)
    return y


@pytest.mark.parametrize("grid", [(20, 14, 11), (48, 48, 48), (130, 7, 5)])
def test_dictionary_spmv_has_the_bits_of_the_value_spmv(ctx, grid):
    A = ctx.generate_hpcg(*grid)
    n = A.info()["n_rows"]
    x = np.random.default_rng(2).uniform(-1.0, 1.0, n)
    y_dict = _spmv(ctx, A, x)
    assert ctx.get_option("spmv_value_bytes") == 1, "HPCG has two distinct values: the dictionary should be in use"
    ctx.set_option("spmv_vdict", 0)
    y_val = _spmv(ctx, A, x)
    assert ctx.get_option("spmv_value_bytes") == 8
    ctx.set_option("spmv_vdict", 1)
    assert np.array_equal(y_dict, y_val)
    rp, col, val = matgen.hpcg(*grid)
    assert np.array_equal(y_dict, port.spmv(rp, col, val, x))
    A.free()


def _banded(n, n_values, seed):
    """pentadiagonal matrix whose entries are drawn from n_values distinct numbers (incl. -0.0 and a subnormal)"""
    rng = np.random.default_rng(seed)
    pool = rng.uniform(-3.0, 3.0, n_values)
    pool[0], pool[1] = -0.0, 5e-324
    I, J = [], []
    for d in (-7, -1, 0, 1, 7):
        r = np.arange(max(0, -d), min(n, n - d))
        I.append(r)
        J.append(r + d)
    I, J = np.concatenate(I), np.concatenate(J)
    key = np.argsort(I * n + J, kind="stable")
    I, J = I[key].astype(np.int32), J[key].astype(np.int32)
    V = pool[rng.integers(0, n_values, I.size)]
    V[: n_values] = pool          # every pool value occurs
    return I, J, V


@pytest.mark.parametrize("n_values,expect_bytes", [(2, 1), (256, 1), (257, 8), (5000, 8)])
def test_dictionary_limit_and_fallback(ctx, n_values, expect_bytes):
    n = 6000
    I, J, V = _banded(n, n_values, 5)
    assert np.unique(V.view(np.uint64)).size == n_values
    A = ctx.upload_coo(n, n, I, J, V, sorted_by_row=True)
    x = np.random.default_rng(6).uniform(-1.0, 1.0, n)
    y = _spmv(ctx, A, x)
    assert ctx.get_option("spmv_value_bytes") == expect_bytes
    rp = np.concatenate([[0], np.cumsum(np.bincount(I, minlength=n))]).astype(np.int32)
    want = port.spmv(rp, J, V, x)
    assert np.array_equal(y.view(np.uint64), want.view(np.uint64))     # bit patterns: -0.0 and subnormal products included
    A.free()


def test_dictionary_solve_has_the_bits_of_the_value_solve(ctx):
    """Whole solves (fused dot / Jacobi / residual epilogues, CUDA-graph replay): identical residual histories."""
    from basic_iterative_solvers_b200 import host
    out = {}
    for on in (1, 0):
        ctx.set_option("spmv_vdict", on)
        out[on] = [host.solve(ctx, m, p, matrix_name="HPCG-24", tol=1e-10, want_x=False).history
                   for m, p in (("cg", "j"), ("bi", "j"), ("j", "none"), ("gm", "none"))]
    ctx.set_option("spmv_vdict", 1)
    for a, b in zip(out[1], out[0]):
        assert np.array_equal(a, b)


def test_dictionary_follows_values_changed_in_place(ctx):
    nx, ny, nz = 16, 12, 9
    A = ctx.generate_hpcg(nx, ny, nz)
    n = A.info()["n_rows"]
    x = np.random.default_rng(3).uniform(-1.0, 1.0, n)
    _spmv(ctx, A, x)                                   # dictionary of the unscaled values is built and used
    assert ctx.get_option("spmv_value_bytes") == 1
    ds = ctx.alloc(n)
    ctx.call("bis_matrix_scale_symmetric", A.h, ds)    # A <- D^-1/2 A D^-1/2: new values (still few distinct ones)
    y = _spmv(ctx, A, x)
    rp, col, val = matgen.hpcg(nx, ny, nz)
    val_s, _ = port.scale_symmetric(rp, col, val)
    assert np.array_equal(y, port.spmv(rp, col, val_s, x))
    ctx.free(ds)
    A.free()
