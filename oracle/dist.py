"""oracle/dist.py -- TEST INFRASTRUCTURE (CPU, numpy + torch.distributed/gloo).

CPU restatement of the multi-GPU plan of basic_iterative_solvers_b200/csrc/bis_dist.cu (new in the
build: the reference is single-process): contiguous row blocks, owned columns renumbered
c - row_begin, ghost columns appended after the owned part in ascending global order (within-row order
of the nonzeros untouched), per-owner receive segments, send-index lists obtained by exchanging the
ghost ids, halo exchange before the boundary rows of every SpMV, allreduce for dot products.
Used by tests/test_partition_cpu.py (world_size-2 gloo) and as the checker of tools/dist_check.py's plan.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import port


NSLAB = 8            # virtual slabs (csrc/bis_internal.cuh: BIS_NSLAB)
RED_CHUNK_MIN = 1024


def partition_rule(n: int):
    """csrc/bis_context.cu bis_partition_rule(): the 8 virtual slabs of an n-row problem and the rows per
    block of a reducing streaming kernel.  Equal row blocks rounded up to whole chunks (whole z-planes
    whenever an eighth of the rows is a whole number of chunks, as for HPCG-128/256/512)."""
    ch = 1
    while ch < (n + 16383) // 16384:
        ch <<= 1
    ch = min(max(ch, RED_CHUNK_MIN), 1 << 20)
    per = (n + NSLAB - 1) // NSLAB
    for unit in (ch, 128):          # whole chunks, else whole SpMV tiles, else exact: never an empty slab
        r = (per + unit - 1) // unit * unit
        if r * (NSLAB - 1) < n:
            per = r
            break
    return [min(n, per * v) for v in range(NSLAB + 1)], ch


def slab(n: int, planes: int, plane: int, rank: int, nranks: int):
    """csrc/bis_context.cu bis_partition_rows(): unions of the virtual slabs when nranks divides 8,
    otherwise contiguous blocks (whole planes per rank when there are enough of them)."""
    if nranks >= 1 and NSLAB % nranks == 0:
        vb, _ = partition_rule(n)
        per = NSLAB // nranks
        return vb[rank * per], vb[(rank + 1) * per]
    if plane > 0 and planes >= nranks:
        q, r = divmod(planes, nranks)
        b = rank * q + min(rank, r)
        e = b + q + (1 if rank < r else 0)
        return b * plane, e * plane
    q, r = divmod(n, nranks)
    b = rank * q + min(rank, r)
    return b, b + q + (1 if rank < r else 0)


class Plan:
    """Halo plan of one rank for local rows [lo, hi) given as CRS with GLOBAL columns."""

    def __init__(self, lo: int, hi: int, n_global: int, rp, col, val):
        self.rank, self.P = dist.get_rank(), dist.get_world_size()
        self.lo, self.hi, self.n = lo, hi, hi - lo
        self.rp, self.val = np.asarray(rp, np.int32), np.asarray(val, np.float64)
        col = np.asarray(col, np.int64)
        bounds = [None] * self.P
        dist.all_gather_object(bounds, (lo, hi))
        self.first = np.array([b[0] for b in bounds] + [bounds[-1][1]], np.int64)
        assert self.first[-1] == n_global and all(bounds[p][1] == bounds[p + 1][0] for p in range(self.P - 1))
        outside = (col < lo) | (col >= hi)
        self.ghost_global = np.unique(col[outside])                    # ascending global ids
        local = np.where(outside, self.n + np.searchsorted(self.ghost_global, col), col - lo)
        self.col = local.astype(np.int32)                               # within-row order untouched
        self.recv_off = np.searchsorted(self.ghost_global, self.first)
        assert self.recv_off[self.rank] == self.recv_off[self.rank + 1]
        # tell every owner which of its rows we need
        wants = [self.ghost_global[self.recv_off[p]:self.recv_off[p + 1]] for p in range(self.P)]
        all_wants = [None] * self.P
        dist.all_gather_object(all_wants, wants)
        self.send_idx = [np.asarray(all_wants[q][self.rank], np.int64) - lo for q in range(self.P)]
        ghost_rows = np.zeros(self.n, bool)
        rows = np.repeat(np.arange(self.n), np.diff(self.rp))
        ghost_rows[rows[outside]] = True
        touched = np.nonzero(ghost_rows)[0]
        low = touched[touched < self.n // 2] if touched.size else touched
        self.interior = (int(low.max()) + 1 if low.size else 0,
                         int(touched[touched >= self.n // 2].min()) if (touched >= self.n // 2).any() else self.n)

    def halo(self, x: np.ndarray) -> np.ndarray:
        ghosts = np.zeros(self.ghost_global.size)
        reqs, bufs = [], {}
        for p in range(self.P):
            if p == self.rank:
                continue
            if self.send_idx[p].size:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x[self.send_idx[p]])), p))
            cnt = self.recv_off[p + 1] - self.recv_off[p]
            if cnt:
                bufs[p] = torch.zeros(int(cnt), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[p], p))
        for r in reqs:
            r.wait()
        for p, b in bufs.items():
            ghosts[self.recv_off[p]:self.recv_off[p + 1]] = b.numpy()
        return ghosts

    def spmv(self, x: np.ndarray) -> np.ndarray:
        return port.spmv(self.rp, self.col, self.val, np.concatenate([x, self.halo(x)]))

    @staticmethod
    def dot(a: np.ndarray, b: np.ndarray) -> float:
        t = torch.tensor([float(np.dot(a, b))], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())


def cg_jacobi(plan: Plan, diag: np.ndarray, b: np.ndarray, x0: np.ndarray, tol: float, max_iters: int):
    """methods/cg.hpp:6-54 + :100-120 with the Jacobi preconditioner, on the partitioned operator."""
    x = x0.copy()
    r = b - plan.spmv(x)
    z = r / diag
    p = z.copy()
    hist = [np.sqrt(plan.dot(r, r))]
    stop = tol * hist[0]
    rz = plan.dot(r, z)
    for _ in range(max_iters):
        ap = plan.spmv(p)
        alpha = rz / plan.dot(ap, p)
        x = x + alpha * p
        r = r - alpha * ap
        z = r / diag
        rz_new = plan.dot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
        hist.append(np.sqrt(plan.dot(r, r)))
        if hist[-1] < stop:
            break
    return x, np.array(hist)
