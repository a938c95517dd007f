"""oracle/refshim.py -- TEST INFRASTRUCTURE.

ctypes binding of oracle/_ref/libbis_ref*.so: the UNMODIFIED reference
(/root/reference) compiled behind the extern "C" shim oracle/ref_harness.cpp.
Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# common.hpp:38-56
PRECOND = {"none": 0, "j": 1, "gs": 2, "bgs": 3, "sgs": 4, "2st": 5, "s2st": 6, "ilu0": 7}
METHOD = {"j": 0, "gs": 1, "sgs": 2, "gm": 3, "cg": 4, "bi": 5}
MAX_ITERS = 1000

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib_path(det=False) -> str:
    """det: False stock flavour, True pinned codegen, "in1" / "in2": -DPRECOND_INNER_ITERS=1 / 2 (two-stage GS)."""
    name = {False: "libbis_ref.so", True: "libbis_ref_det.so", "in1": "libbis_ref_in1.so", "in2": "libbis_ref_in2.so"}[det]
    return os.path.join(HERE, "_ref", name)


def available(det=False) -> bool:
    return os.path.exists(lib_path(det))


_libs: dict = {}


def load(det=False) -> C.CDLL:
    if det in _libs:
        return _libs[det]
    lib = C.CDLL(lib_path(det))
    crs = [C.c_int, C.c_int, _ip, _ip, _dp]
    lib.ref_omp_max_threads.restype = C.c_int
    lib.ref_omp_set_threads.argtypes = [C.c_int]
    lib.ref_set_max_iters.argtypes = [C.c_int]
    lib.ref_spmv.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
    lib.ref_sptrsv.argtypes = crs + [_dp, _dp, _dp]
    lib.ref_bsptrsv.argtypes = crs + [_dp, _dp, _dp]
    for f in ("ref_subtract_vectors", "ref_sum_vectors", "ref_elemwise_mult_vectors",
              "ref_elemwise_div_vectors"):
        getattr(lib, f).argtypes = [_dp, _dp, _dp, C.c_int, C.c_double]
    lib.ref_scale.argtypes = [_dp, _dp, C.c_double, C.c_int]
    lib.ref_copy_vector.argtypes = [_dp, _dp, C.c_int]
    lib.ref_dot.argtypes = [_dp, _dp, C.c_int]
    lib.ref_dot.restype = C.c_double
    lib.ref_euclidean_vec_norm.argtypes = [_dp, C.c_int]
    lib.ref_euclidean_vec_norm.restype = C.c_double
    lib.ref_normalize_x.argtypes = [_dp, _dp, _dp, _dp, C.c_int]
    lib.ref_compute_residual.argtypes = crs + [_dp, _dp, _dp, _dp]
    lib.ref_apply_preconditioner.argtypes = (
        [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, _ip, _ip, _dp] + [_dp] * 8)
    lib.ref_gmres_least_squares.argtypes = [C.c_int] * 3 + [_dp] * 6
    lib.ref_gmres_update_g.argtypes = [C.c_int] * 3 + [_dp] * 3 + [C.c_double]
    lib.ref_gmres_update_g.restype = C.c_double
    lib.ref_factor_begin.argtypes = crs + [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_factor_fetch.argtypes = [_ip, _ip, _dp, _ip, _ip, _dp, _dp, _dp, _dp, _dp]
    lib.ref_solve.argtypes = crs + [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_int,
                                                   _dp, _dp, _dp, _ip, _dp]
    lib.ref_solve.restype = C.c_int
    lib.ref_read_mtx_begin.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_read_mtx_fetch.argtypes = [_ip, _ip, _dp]
    _libs[det] = lib
    return lib


def _crs(rp, col, val):
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    if col.size == 0:  # ndpointer rejects nothing, but keep a valid address
        col = np.zeros(1, np.int32)
        val = np.zeros(1, np.float64)
    return rp, col, val


def spmv(rp, col, val, x, n_cols=None, det=False):
    lib = load(det)
    rp, col, val = _crs(rp, col, val)
    n = rp.size - 1
    y = np.zeros(n)
    lib.ref_spmv(n, n if n_cols is None else n_cols, int(rp[-1]), rp, col, val,
                 np.ascontiguousarray(x, dtype=np.float64), y)
    return y


def sptrsv(rp, col, val, D, b, backward=False, x_init=None, det=False):
    lib = load(det)
    rp, col, val = _crs(rp, col, val)
    n = rp.size - 1
    x = np.zeros(n) if x_init is None else np.array(x_init, dtype=np.float64)
    f = lib.ref_bsptrsv if backward else lib.ref_sptrsv
    f(n, int(rp[-1]), rp, col, val, x, np.ascontiguousarray(D), np.ascontiguousarray(b))
    return x


def sptrsv_inplace(rp, col, val, D, xb, backward=False, det=False):
    """x aliases b (gmres.hpp:288-291, bicgstab.hpp:157-160)."""
    lib = load(det)
    rp, col, val = _crs(rp, col, val)
    n = rp.size - 1
    x = np.array(xb, dtype=np.float64)
    f = lib.ref_bsptrsv if backward else lib.ref_sptrsv
    f(n, int(rp[-1]), rp, col, val, x, np.ascontiguousarray(D), x)
    return x


@dataclass
class Factors:
    l_rp: np.ndarray
    l_col: np.ndarray
    l_val: np.ndarray
    u_rp: np.ndarray
    u_col: np.ndarray
    u_val: np.ndarray
    A_D: np.ndarray
    A_D_inv: np.ndarray
    L_D: np.ndarray
    U_D: np.ndarray


def factor(rp, col, val, precond="none", ilu0_old=True, det=False) -> Factors:
    lib = load(det)
    rp, col, val = _crs(rp, col, val)
    n = rp.size - 1
    nl, nu = C.c_int(0), C.c_int(0)
    lib.ref_factor_begin(n, int(rp[-1]), rp, col, val, PRECOND[precond], int(ilu0_old),
                         C.byref(nl), C.byref(nu))
    f = Factors(np.zeros(n + 1, np.int32), np.zeros(max(nl.value, 1), np.int32),
                np.zeros(max(nl.value, 1)), np.zeros(n + 1, np.int32),
                np.zeros(max(nu.value, 1), np.int32), np.zeros(max(nu.value, 1)),
                np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n))
    lib.ref_factor_fetch(f.l_rp, f.l_col, f.l_val, f.u_rp, f.u_col, f.u_val,
                         f.A_D, f.A_D_inv, f.L_D, f.U_D)
    f.l_col, f.l_val = f.l_col[:nl.value], f.l_val[:nl.value]
    f.u_col, f.u_val = f.u_col[:nu.value], f.u_val[:nu.value]
    return f


def apply_preconditioner(precond, fac: Factors, inp, inplace=False, det=False):
    lib = load(det)
    n = fac.A_D.size
    l = _crs(fac.l_rp, fac.l_col, fac.l_val)
    u = _crs(fac.u_rp, fac.u_col, fac.u_val)
    inp = np.array(inp, dtype=np.float64)
    out = inp if inplace else np.zeros(n)
    tmp, work = np.zeros(n), np.zeros(n)
    lib.ref_apply_preconditioner(PRECOND[precond], n, int(l[0][-1]), *l, int(u[0][-1]), *u,
                                 fac.A_D.copy(), fac.A_D_inv.copy(), fac.L_D.copy(),
                                 fac.U_D.copy(), out, inp, tmp, work)
    return out


@dataclass
class SolveResult:
    history: np.ndarray          # collected_residual_norms[0:count]
    final_true_residual: float   # ||b - A x_star|| from save_x_star
    iter_count: int
    converged: bool
    restarts: int
    stopping_criteria: float
    x_star: np.ndarray
    iter_time: np.ndarray
    iterate_time: float
    spmv_time: float
    precond_time: float
    solve_time: float


def solve(rp, col, val, method, precond="none", restart_len=10, num_scale=False,
          ilu0_old=True, use_ref_preprocessing=False, b=None, x0=None, quiet=True,
          det=False, threads=None) -> SolveResult:
    lib = load(det)
    if threads is not None:
        lib.ref_omp_set_threads(int(threads))
    rp, col, val = _crs(rp, col, val)
    n = rp.size - 1
    hist = np.zeros(2 * MAX_ITERS)
    itime = np.zeros(2 * MAX_ITERS)
    xs = np.zeros(n)
    oi = np.zeros(4, np.int32)
    od = np.zeros(6)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    xx = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    rc = lib.ref_solve(n, int(rp[-1]), rp, col, val, METHOD[method], PRECOND[precond],
                       restart_len, int(num_scale), int(ilu0_old), int(use_ref_preprocessing),
                       None if bb is None else bb.ctypes.data, None if xx is None else xx.ctypes.data,
                       int(quiet), hist, itime, xs, oi, od)
    if rc != 0:
        raise RuntimeError(f"ref_solve failed rc={rc}")
    cnt = int(oi[1])
    return SolveResult(hist[:cnt].copy(), float(od[1]), int(oi[0]), bool(oi[2]), int(oi[3]),
                       float(od[0]), xs, itime[:cnt + 2].copy(), float(od[2]), float(od[3]),
                       float(od[4]), float(od[5]))


def read_mtx(path: str, det=False):
    lib = load(det)
    n, nnz = C.c_int(0), C.c_int(0)
    if lib.ref_read_mtx_begin(path.encode(), C.byref(n), C.byref(nnz)) != 0:
        raise RuntimeError("ref_read_mtx failed")
    rp = np.zeros(n.value + 1, np.int32)
    col = np.zeros(nnz.value, np.int32)
    val = np.zeros(nnz.value)
    lib.ref_read_mtx_fetch(rp, col, val)
    return rp, col, val
