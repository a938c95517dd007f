"""world_size-2 (and 3) gloo tests of the N>1 plan on CPU: row partition, ghost renumbering, halo index
lists, halo exchange, allreduce (oracle/dist.py restates csrc/bis_dist.cu; the CUDA path itself is checked
on >= 2 GPUs by tools/dist_check.py).  Properties checked:
 - bench.py's slab_rows == the library's slab rule; blocks are contiguous and cover all rows;
 - the partitioned SpMV equals the rows of the global SpMV BIT FOR BIT (ghost renumbering keeps the
   within-row order, so the summation order is partition-independent);
 - the interior row range touches no ghost column;
 - Jacobi-preconditioned CG on the partitioned operator follows the single-process oracle's residual
   history to 1e-10 * ||r0|| with the same iteration count (tol 1e-10).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, dims, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import dist as odist, matgen, port
        import bench
        nx, ny, nz = dims
        n = nx * ny * nz
        lo, hi = odist.slab(n, nz, nx * ny, rank, world)
        assert (lo, hi) == bench.slab_rows(n, nz, nx * ny, rank, world)
        rp, col, val = matgen.hpcg(nx, ny, nz, row_begin=lo, row_end=hi)
        plan = odist.Plan(lo, hi, n, rp, col, val)
        grp, gcol, gval = matgen.hpcg(nx, ny, nz)
        rng = np.random.default_rng(17)
        xg = rng.uniform(-1, 1, n)
        y_local = plan.spmv(xg[lo:hi])
        y_global = port.spmv(grp, gcol, gval, xg)
        assert np.array_equal(y_local, y_global[lo:hi]), "partitioned SpMV differs from the global one"
        ib, ie = plan.interior
        rows = np.repeat(np.arange(plan.n), np.diff(plan.rp))
        ghost_rows = np.unique(rows[plan.col >= plan.n])
        assert not ((ghost_rows >= ib) & (ghost_rows < ie)).any()
        # z-slab halo: one plane per neighbour
        plane_aligned = nz >= world and lo % (nx * ny) == 0 and hi % (nx * ny) == 0
        expect = (nx * ny) * ((1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)) if plane_aligned else None
        if expect is not None:
            assert plan.ghost_global.size == expect
        diag = np.full(hi - lo, 26.0)
        b = rng.uniform(0.5, 1.5, n)
        x0 = np.full(n, 0.1)
        x, hist = odist.cg_jacobi(plan, diag, b[lo:hi], x0[lo:hi], 1e-10, 500)
        want = port.solve(grp, gcol, gval, "cg", "j", b=b, x0=x0, tol=1e-10)
        assert len(hist) - 1 == want.iter_count, (len(hist) - 1, want.iter_count)
        assert np.max(np.abs(hist - want.history)) <= 1e-10 * want.history[0]
        assert np.max(np.abs(x - want.x_star[lo:hi])) <= 1e-9
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (10, 9, 8)), (3, (8, 8, 7)), (2, (6, 5, 1))])
def test_partition_plan_gloo(world, dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
