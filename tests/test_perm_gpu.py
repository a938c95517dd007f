"""The permutation seam as a labelled mode (option "perm_mode" = 1 = the reference's PERM_MODE = C, multicolouring;
preprocessing.hpp:52-65, utilities/smax_helpers.hpp:44-80 -- SMAX itself is absent, so the oracle of this mode is the
CPU restatement run on the explicitly permuted system): the device's permutation, the permuted matrix and whole
solves, against numpy / oracle/port on the same permutation."""
import numpy as np
import pytest

from conftest import HIST_TOL
from oracle import matgen, port

from basic_iterative_solvers_b200 import host

pytestmark = pytest.mark.gpu


def permute_crs(rp, col, val, perm, inv):
    """P A P^T with the stored order kept inside a row (csrc/bis_perm.cu: perm_fill_kernel)."""
    rp = np.asarray(rp, np.int64)
    lens = np.diff(rp)[perm]
    rp2 = np.concatenate([[0], np.cumsum(lens)])
    idx = np.concatenate([np.arange(rp[o], rp[o + 1]) for o in perm]) if perm.size else np.zeros(0, np.int64)
    return rp2.astype(np.int32), inv[np.asarray(col)[idx]].astype(np.int32), np.asarray(val)[idx]


def test_grid_colouring_is_the_eight_colour_rule(ctx):
    nx, ny, nz = 10, 8, 6
    A = ctx.generate_hpcg(nx, ny, nz)
    perm, inv, colours = ctx.colouring_permutation(A)
    n = nx * ny * nz
    r = np.arange(n)
    colour = (r % nx & 1) | ((r // nx % ny & 1) << 1) | ((r // (nx * ny) & 1) << 2)
    want = np.argsort(colour, kind="stable").astype(np.int32)
    assert colours == 8 and np.array_equal(perm, want)
    assert np.array_equal(inv[perm], np.arange(n))
    A.free()


@pytest.mark.parametrize("method,pre", [("cg", "sgs"), ("gm", "ilu0"), ("bi", "gs"), ("sgs", "none"), ("cg", "j")])
def test_permuted_solve_matches_oracle_on_permuted_system(ctx, method, pre):
    nx, ny, nz = 12, 9, 7
    rp, col, val = matgen.hpcg(nx, ny, nz)
    n = rp.size - 1
    A = ctx.generate_hpcg(nx, ny, nz)
    perm, inv, colours = ctx.colouring_permutation(A)
    A.free()
    rng = np.random.default_rng(8)
    b, x0 = rng.uniform(0.5, 1.5, n), rng.uniform(-0.1, 0.1, n)
    prp, pcol, pval = permute_crs(rp, col, val, perm, inv)
    # the reference permutes b and x_0 AFTER init_structs has copied x_0 into the solver's iterate
    # (preprocessing.hpp:33 vs :60): the solve starts from the unpermuted copy
    want = port.solve(prp, pcol, pval, method, pre, b=b[perm], x0=x0, tol=1e-10)
    ctx.set_option("perm_mode", 1)
    try:
        got = host.solve(ctx, method, pre, matrix_name=f"HPCG-{nx}-{ny}-{nz}", b=b, x0=x0, tol=1e-10)
    finally:
        ctx.set_option("perm_mode", 0)
    k = min(got.history.size, want.history.size)
    assert np.max(np.abs(got.history[:k] - want.history[:k])) <= HIST_TOL * want.history[0]
    assert got.iter_count == want.iter_count and got.converged == want.converged
    plain = host.solve(ctx, method, pre, matrix_name=f"HPCG-{nx}-{ny}-{nz}", b=b, x0=x0, tol=1e-10, want_x=False)
    assert plain.history.size != got.history.size or not np.array_equal(plain.history, got.history) or pre == "j"


def test_unstructured_colouring_and_level_count(ctx):
    """No grid hint: Luby / Jones-Plassmann rounds.  The classes must be independent sets, so the strict lower
    factor of the permuted matrix has at most as many levels as there are colours."""
    rng = np.random.default_rng(12)
    n = 600
    rows, cols = [], []
    for r in range(n):
        for c in rng.integers(0, n, size=4):
            if c != r:
                rows += [r, int(c)]
                cols += [int(c), r]
    rows += list(range(n))
    cols += list(range(n))
    key = np.unique(np.array(rows, np.int64) * n + np.array(cols, np.int64))
    I, J = (key // n).astype(np.int32), (key % n).astype(np.int32)
    V = np.where(I == J, 40.0, -1.0)
    A = ctx.upload_coo(n, n, I, J, V, sorted_by_row=True)
    perm, inv, colours = ctx.colouring_permutation(A)
    assert np.array_equal(np.sort(perm), np.arange(n)) and 2 <= colours <= 64
    rp, col, val = A.download()
    prp, pcol, pval = permute_crs(rp, col, val, perm, inv)
    B = ctx.upload_crs(prp, pcol, pval)
    L, U = ctx.split_triangular(B)
    ctx.set_option("trsv_variant", 3)
    try:
        N = n
        D, bb, x = ctx.upload(np.full(N, 40.0)), ctx.upload(np.ones(N)), ctx.alloc(N)
        ctx.call("bis_sptrsv", L.h, x, D, bb)      # builds the level sets
        ctx.sync()
    finally:
        ctx.set_option("trsv_variant", 0)
    assert 1 <= L.info()["n_levels"] <= colours
    for m in (L, U, B, A):
        m.free()


# ---- breadth-first family (perm_mode 2 = BFS levels, 3 = reverse Cuthill-McKee, 4 = Cuthill-McKee) ------------------
def bfs_orders(rp, col):
    """numpy restatement of csrc/bis_perm.cu: bis_matrix_bfs_permutation -> (bfs, cm, n_levels)."""
    n = rp.size - 1
    rows = [np.array([c for c in col[rp[r]:rp[r + 1]] if c != r], np.int64) for r in range(n)]
    deg = np.array([r.size for r in rows])
    level = np.full(n, -1)
    cur = 0
    while (level < 0).any():
        un = np.nonzero(level < 0)[0]
        root = un[np.lexsort((un, deg[un]))[0]]
        level[root] = cur
        while True:
            found = False
            for r in np.nonzero(level == cur)[0]:
                for c in rows[r]:
                    if level[c] < 0:
                        level[c] = cur + 1
                        found = True
            cur += 1
            if not found:
                break
    bfs = np.lexsort((np.arange(n), level))
    pos = np.zeros(n, np.int64)
    cm = []
    for l in range(cur):
        ids = np.nonzero(level == l)[0]
        parent = np.array([min([pos[c] for c in rows[v] if level[c] == l - 1], default=0xFFFFFFFF) for v in ids], np.int64)
        ids = ids[np.lexsort((ids, np.minimum(deg[ids], 0xFFFFFF), parent))]
        for v in ids:
            pos[v] = len(cm)
            cm.append(v)
    return bfs.astype(np.int32), np.array(cm, np.int32), cur


def _random_symmetric(ctx, n, seed, fragments=1):
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    size = n // fragments
    for r in range(n):
        lo = (r // size) * size
        hi = min(n, lo + size) if r // size < fragments - 1 else n
        for c in rng.integers(lo, hi, size=3):
            if c != r:
                rows += [r, int(c)]
                cols += [int(c), r]
    rows += list(range(n))
    cols += list(range(n))
    key = np.unique(np.array(rows, np.int64) * n + np.array(cols, np.int64))
    I, J = (key // n).astype(np.int32), (key % n).astype(np.int32)
    return ctx.upload_coo(n, n, I, J, np.where(I == J, 40.0, -1.0), sorted_by_row=True)


@pytest.mark.parametrize("case", ["grid", "random", "fragments"])
def test_bfs_and_cuthill_mckee_permutations_match_numpy(ctx, case):
    if case == "grid":
        A = ctx.generate_hpcg(9, 7, 5)
    elif case == "random":
        A = _random_symmetric(ctx, 700, 21)
    else:
        A = _random_symmetric(ctx, 900, 22, fragments=3)      # three components: the search restarts twice
    rp, col, _ = A.download()
    n = rp.size - 1
    bfs, cm, levels = bfs_orders(np.asarray(rp, np.int64), np.asarray(col, np.int64))
    for mode, want in ((2, bfs), (4, cm), (3, cm[::-1])):
        perm, inv, nl = ctx.bfs_permutation(A, mode)
        assert nl == levels
        assert np.array_equal(perm, want), f"mode {mode}"
        assert np.array_equal(inv[perm], np.arange(n))
    A.free()


def test_reverse_cuthill_mckee_shrinks_the_bandwidth(ctx):
    """What the ordering is for: a randomly renumbered 2-D grid gets a bandwidth of the order of its side back."""
    nx, ny = 30, 20
    n = nx * ny
    rng = np.random.default_rng(4)
    shuffle = rng.permutation(n)
    I, J = [], []
    for y in range(ny):
        for x in range(nx):
            r = y * nx + x
            for dx, dy in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)):
                if 0 <= x + dx < nx and 0 <= y + dy < ny:
                    I.append(shuffle[r])
                    J.append(shuffle[(y + dy) * nx + x + dx])
    key = np.unique(np.array(I, np.int64) * n + np.array(J, np.int64))
    I, J = (key // n).astype(np.int32), (key % n).astype(np.int32)
    A = ctx.upload_coo(n, n, I, J, np.where(I == J, 4.0, -1.0), sorted_by_row=True)
    perm, inv, _ = ctx.bfs_permutation(A, 3)
    before = int(np.max(np.abs(I.astype(np.int64) - J)))
    after = int(np.max(np.abs(inv[I].astype(np.int64) - inv[J])))
    assert before > 10 * nx and after <= 2 * nx
    A.free()


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("method,pre", [("cg", "sgs"), ("gm", "ilu0")])
def test_bfs_permuted_solve_matches_oracle_on_permuted_system(ctx, mode, method, pre):
    nx, ny, nz = 11, 8, 6
    rp, col, val = matgen.hpcg(nx, ny, nz)
    n = rp.size - 1
    A = ctx.generate_hpcg(nx, ny, nz)
    perm, inv, _ = ctx.bfs_permutation(A, mode)
    A.free()
    rng = np.random.default_rng(8)
    b, x0 = rng.uniform(0.5, 1.5, n), rng.uniform(-0.1, 0.1, n)
    prp, pcol, pval = permute_crs(rp, col, val, perm, inv)
    want = port.solve(prp, pcol, pval, method, pre, b=b[perm], x0=x0, tol=1e-10)
    ctx.set_option("perm_mode", mode)
    try:
        got = host.solve(ctx, method, pre, matrix_name=f"HPCG-{nx}-{ny}-{nz}", b=b, x0=x0, tol=1e-10)
    finally:
        ctx.set_option("perm_mode", 0)
    k = min(got.history.size, want.history.size)
    assert np.max(np.abs(got.history[:k] - want.history[:k])) <= HIST_TOL * want.history[0]
    assert got.iter_count == want.iter_count and got.converged == want.converged
