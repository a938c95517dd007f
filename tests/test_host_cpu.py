"""CPU tests of the host logic and of the C-ABI surface (no compute calls: there is no GPU here).

 - libbis_b200.so loads and exports every symbol include/bis_b200.h declares;
 - without a GPU the library fails loudly (no CPU fallback);
 - host/ preprocessing (split, diagonal peel, ILU(0)) is bit-identical to the compiled
   reference's outputs (tests/golden) and to the oracle;
 - host/ matrix generators, MatrixMarket reader, CLI grammar, GMRES Givens update.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden
from oracle import matgen, port

from basic_iterative_solvers_b200 import capi, host


def test_library_exports_every_declared_symbol(built):
    lib = capi.load()
    names = capi.declared_symbols()
    assert len(names) >= 55
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes table covers the whole header too
    assert sorted(set(names) - set(capi._SIGS) - {"bis_last_error"}) == []
    assert lib.bis_version() == 100


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BisError):
        capi.Context(0)
    rp, col, val = matgen.hpcg(4)
    with pytest.raises(Exception):
        host.solve(None, "cg", crs=(rp, col, val))


def test_parse_cli_grammar(built):
    a = host.parse_cli(["bis", "HPCG-16", "-cg", "-p", "sgs"])
    assert a == {"method": "cg", "precond": "sgs", "restart_length": 10, "num_scale": False}
    a = host.parse_cli(["bis", "m.mtx", "-gm", "-p", "ilu0", "-rl", "25"])
    assert a["method"] == "gm" and a["precond"] == "ilu0" and a["restart_length"] == 25
    for m in ("j", "gs", "sgs", "cg", "gm", "bi"):
        assert host.parse_cli(["bis", "x", "-" + m])["method"] == m
    for p in ("j", "gs", "bgs", "sgs", "2st", "s2st", "ilu0"):
        assert host.parse_cli(["bis", "x", "-cg", "-p", p])["precond"] == p
    # reference behaviour: unknown method / preconditioner / missing arguments are fatal
    for bad in (["bis", "x"], ["bis", "x", "-xx"], ["bis", "x", "-cg", "-p", "ilut"], ["bis", "x", "-cg", "-p"]):
        with pytest.raises(capi.BisError):
            host.parse_cli(bad)


@pytest.mark.parametrize("name", ["fdm2d16", "band_klein"])
def test_host_factor_matches_reference(built, name):
    g = golden(name)
    rp, col, val = g["rp"], g["col"], g["val"]
    f = host.factor(rp, col, val, "sgs")
    for k in ("l_rp", "l_col", "l_val", "u_rp", "u_col", "u_val", "A_D", "A_D_inv"):
        assert np.array_equal(getattr(f, k), g["k__split__" + k]), k
    f = host.factor(rp, col, val, "ilu0")
    for k in ("l_rp", "l_col", "l_val", "u_rp", "u_col", "u_val", "L_D", "U_D"):
        assert np.array_equal(getattr(f, k), g["k__ilu0__" + k]), k


def test_host_factor_matches_oracle_hpcg_and_anderson(built):
    for rp, col, val in (matgen.hpcg(9, 7, 5), matgen.anderson(7, 6, 5, 5.0, 1.0, 3, True)):
        rp = rp.astype(np.int32)
        for pre in ("sgs", "ilu0"):
            a, b = host.factor(rp, col, val, pre), port.factor(rp, col, val, pre)
            for k in ("l_rp", "l_col", "l_val", "u_rp", "u_col", "u_val", "A_D", "A_D_inv", "L_D", "U_D"):
                assert np.array_equal(getattr(a, k), getattr(b, k)), (pre, k)


def test_host_generators_match_matgen(built):
    for name, ref in (("HPCG-6", matgen.hpcg(6)), ("HPCG-5-4-3", matgen.hpcg(5, 4, 3)),
                      ("Anderson,Lx=6,Ly=5,Lz=4,ranpot=5.0", matgen.anderson(6, 5, 4, 5.0, 1.0, 1, False))):
        rp, col, val = host.matrix(name)
        assert np.array_equal(rp, ref[0]) and np.array_equal(col, ref[1]) and np.array_equal(val, ref[2]), name


def test_host_mtx_reader_keeps_file_order(built, tmp_path):
    # general file, rows shuffled, columns of one row in non-ascending file order: the reader
    # stable-sorts by row only (sparse_matrix.hpp:332-344)
    p = tmp_path / "m.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% c\n3 3 5\n"
                 "2 3 6.0\n1 1 1.5\n2 1 4.0\n3 3 9.0\n2 2 5.0\n")
    rp, col, val = host.matrix(str(p))
    assert rp.tolist() == [0, 1, 4, 5]
    assert col.tolist() == [0, 2, 0, 1, 2] and val.tolist() == [1.5, 6.0, 4.0, 5.0, 9.0]
    # symmetric file: each off-diagonal entry is mirrored
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n3 3 4\n1 1 2\n2 1 -1\n2 2 2\n3 3 2\n")
    rp, col, val = host.matrix(str(p))
    d = np.zeros((3, 3))
    for r in range(3):
        for k in range(rp[r], rp[r + 1]):
            d[r, col[k]] = val[k]
    assert np.array_equal(d, np.array([[2, -1, 0], [-1, 2, 0], [0, 0, 2.0]]))
    with pytest.raises(capi.BisError):
        host.matrix(str(tmp_path / "missing.mtx"))


def test_host_mtx_reader_matches_reference_reader(built, tmp_path):
    # fdm2d16.npz holds the CRS the reference's own reader produced from FDM-2d-16.mtx
    # (symmetric storage).  Re-emit the lower triangle in the reference's order and read it back.
    g = golden("fdm2d16")
    rp, col, val = g["rp"], g["col"], g["val"]
    n = rp.size - 1
    rows = np.repeat(np.arange(n), np.diff(rp))
    dense = np.zeros((n, n))
    dense[rows, col] = val
    assert np.array_equal(dense, dense.T)
    lines = [f"{r + 1} {c + 1} {dense[r, c]:.17g}" for c in range(n) for r in range(c, n) if dense[r, c] != 0.0]
    p = tmp_path / "fdm.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n" + f"{n} {n} {len(lines)}\n" + "\n".join(lines) + "\n")
    rp2, col2, val2 = host.matrix(str(p))
    assert np.array_equal(rp2, rp)
    d2 = np.zeros((n, n))
    d2[np.repeat(np.arange(n), np.diff(rp2)), col2] = val2
    assert np.array_equal(d2, dense)


def test_host_gmres_givens_matches_oracle(built):
    rng = np.random.default_rng(5)
    m = 6
    lib = port.load()
    H = np.zeros((m + 1) * m)
    state_h = [np.zeros((m + 1) * (m + 1)), H.copy(), np.zeros((m + 1) * m), np.eye(m + 1).ravel().copy(),
               np.eye(m + 1).ravel().copy(), np.zeros((m + 1) * m)]
    state_o = [a.copy() for a in state_h]
    g_h, gt_h = np.zeros(m + 1), np.zeros(m + 1)
    g_o, gt_o = np.zeros(m + 1), np.zeros(m + 1)
    beta = 3.25
    for k in range(m):
        colv = rng.uniform(-1, 1, k + 2)
        for st in (state_h, state_o):
            for j in range(k + 2):
                st[1][j * m + k] = colv[j]
        host.gmres_least_squares(k, m, *state_h)
        lib.o_gmres_least_squares(k, m, *state_o)
        for a, b in zip(state_h, state_o):
            assert np.array_equal(a, b)
        rh = host.gmres_update_g(k, m, state_h[3], g_h, gt_h, beta)
        ro = lib.o_gmres_update_g(k, m, state_o[3], g_o, gt_o, beta)
        assert rh == ro and np.array_equal(g_h, g_o)


def test_harness_run_ahead_order_and_stop_decisions():
    """solve() / harness_step (host/solver_harness.hpp) on a device-free recording solver: with run-ahead the
    next iterate() is enqueued between the norm readback's begin and end, the decisions are those of the
    reference's order (solver_harness.hpp:17-50), and nothing is enqueued past max_iters."""
    # reference order: iterate, sample, exchange, check_restart
    t, r = host.harness_trace(False, 50, 1e-3, 0.5)
    assert r == {"iter_count": 10, "iterates": 10, "converged": True, "history": 11}      # 0.5^10 < 1e-3
    assert t == "IsXR|" * 10
    # run-ahead: one iterate in flight behind every readback; the last one is the discarded speculation
    t, r = host.harness_trace(True, 50, 1e-3, 0.5)
    assert r == {"iter_count": 10, "iterates": 11, "converged": True, "history": 11}
    assert t == "IbXIeR|" + "bXIeR|" * 9
    # does not converge: exactly max_iters iterates, the last pass does not speculate
    t, r = host.harness_trace(True, 7, 1e-30, 0.9)
    assert r == {"iter_count": 7, "iterates": 7, "converged": False, "history": 8}
    assert t == "IbXIeR|" + "bXIeR|" * 5 + "bXeR|"
    t, r = host.harness_trace(False, 7, 1e-30, 0.9)
    assert r["iterates"] == 7 and t == "IsXR|" * 7


def test_integration_stub_compiles_against_the_header(tmp_path):
    """The reference-side binding shown in INTEGRATION.md section 2 is real code: it must compile against
    include/bis_b200.h (g++ -fsyntax-only; no CUDA, no device)."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    m = re.search(r"```cpp\n(.*?)```", text, re.S)
    assert m, "INTEGRATION.md lost its C++ stub"
    src = tmp_path / "bis_b200_binding.cpp"
    src.write_text(m.group(1) + "\nint main() { return 0; }\n")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(root, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
