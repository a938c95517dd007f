"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank solves the row-partitioned problem; rank 0 also solves the whole problem on its own GPU
(single-GPU context) and compares.  With the peer-memory transport and a rank count that divides 8 the
reductions are partition-invariant (8 fixed slab sums added in a fixed order, csrc/bis_internal.cuh): EVERY
history -- CG, BiCGSTAB, GMRES, Jacobi -- must be bit-identical to the single-GPU one, iteration counts
included.  The NCCL fallback transport adds per-rank sums instead: compared to 1e-10 * ||r0||."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi, host  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.frombuffer(bytearray(capi.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(idt, 0)
ctx = capi.Context(local, rank, world, bytes(idt.cpu().numpy().tobytes()))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
name = f"HPCG-{n}"
ok = True


def cmp_len(method, a, b):
    """BiCGSTAB amplifies rounding differences: the compiled reference itself takes 134 / 131 / 121
    iterations on HPCG-96 `-bi -p j` at 1 / 8 / 4 OpenMP threads and its histories differ by 0.25 * r0
    (4.5e-6 * r0 already within the first 30 entries; measured with oracle/_ref).  Only the ORDER in which
    the ranks' partial sums are added separates the runs compared here, so BiCGSTAB's first 10 residuals
    are held to 1e-10 * r0 and the rest to the iteration count; every other method is compared over the
    whole history."""
    return min(a, b, 10) if method == "bi" else min(a, b)


results = {}
peer = ctx.info()["peer_memory"]
if rank == 0:
    print(f"transport: {'peer memory (CUDA IPC + NVLink stores in the kernels)' if peer else 'NCCL'}", flush=True)
for method, pre in (("cg", "none"), ("cg", "j"), ("bi", "j"), ("j", "none"), ("gm", "j")):
    r = host.solve(ctx, method, pre, matrix_name=name, want_x=False, max_iters=300)
    results[(method, pre)] = r
dist.barrier()
if peer:
    # the same solves over NCCL: only the order in which the ranks' partial sums are added differs
    ctx.set_option("dist_p2p", 0)
    dist.barrier()
    for (method, pre), r in list(results.items()):
        q = host.solve(ctx, method, pre, matrix_name=name, want_x=False, max_iters=300)
        k = cmp_len(method, q.history.size, r.history.size)
        err = float(np.max(np.abs(q.history[:k] - r.history[:k])) / r.history[0])
        if rank == 0:
            print(f"{name} -{method} -p {pre}: peer-memory vs NCCL transport: its {r.iter_count} vs {q.iter_count}, "
                  f"max |dr|/r0 = {err:.2e}", flush=True)
        ok &= err <= 1e-10
    ctx.set_option("dist_p2p", 1)
    dist.barrier()
    # peer-memory transport with separate pack / interior / wait / strip launches instead of the fused kernel
    ctx.set_option("spmv_fused", 0)
    dist.barrier()
    for (method, pre), r in list(results.items()):
        q = host.solve(ctx, method, pre, matrix_name=name, want_x=False, max_iters=300)
        k = cmp_len(method, q.history.size, r.history.size)
        err = float(np.max(np.abs(q.history[:k] - r.history[:k])) / r.history[0])
        if rank == 0:
            print(f"{name} -{method} -p {pre}: fused SpMV kernel vs separate launches: its {r.iter_count} vs {q.iter_count}, "
                  f"max |dr|/r0 = {err:.2e}", flush=True)
        ok &= np.array_equal(q.history, r.history)   # same kernels, same sums: bit-identical
    ctx.set_option("spmv_fused", 1)
    dist.barrier()

# ---- unstructured matrix: every rank needs ghosts from every other rank (SpMV variant 2 with the halo) --------
rng = np.random.default_rng(5)
ng = 40000
lens = rng.integers(1, 12, size=ng)
rows = np.repeat(np.arange(ng), lens)
cols = rng.integers(0, ng, size=rows.size)
keep = np.ones(rows.size, bool)
keep[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])          # no duplicate neighbours in storage order
rows, cols = rows[keep], cols[keep]
vals = rng.uniform(-1.0, 1.0, size=rows.size)
grp = np.zeros(ng + 1, np.int64)
np.add.at(grp, rows + 1, 1)
grp = np.cumsum(grp)
xg = np.sin(np.arange(ng) * 0.37)
lo, hi = rank * ng // world, (rank + 1) * ng // world
Ad = ctx.upload_crs_distributed(lo, ng, grp[lo:hi + 1] - grp[lo], cols[grp[lo]:grp[hi]], vals[grp[lo]:grp[hi]])
dxl, dyl, dzl = ctx.upload(xg[lo:hi]), ctx.alloc(hi - lo), ctx.alloc(hi - lo)
got = {}
for transport in ((1, 0) if peer else (0,)):
    ctx.set_option("dist_p2p", transport)
    dist.barrier()
    ctx.call("bis_spmv", Ad.h, dxl, dyl)          # y = A x
    ctx.call("bis_spmv", Ad.h, dyl, dzl)          # z = A y: two exchanges in a row, no reduction between them
    ctx.call("bis_spmv", Ad.h, dzl, dyl)          # y = A z: third exchange reuses the first ghost copy
    ctx.call("bis_spmv_dot", Ad.h, dxl, dzl, dxl, 40, 41)
    yl = torch.from_numpy(ctx.download(dyl, hi - lo)).cuda()
    parts = [torch.empty((r + 1) * ng // world - r * ng // world, dtype=torch.float64, device="cuda") for r in range(world)]
    dist.all_gather(parts, yl)
    got[transport] = (torch.cat(parts).cpu().numpy(), ctx.scalars(40, 2).copy())
if peer:
    ctx.set_option("dist_p2p", 1)
dist.barrier()
if rank == 0:
    with capi.Context(local) as solo:
        A1 = solo.upload_crs(grp.astype(np.int32), cols.astype(np.int32), vals)
        a, b_, c_ = solo.upload(xg), solo.alloc(ng), solo.alloc(ng)
        solo.call("bis_spmv", A1.h, a, b_)
        solo.call("bis_spmv", A1.h, b_, c_)
        solo.call("bis_spmv", A1.h, c_, b_)
        solo.call("bis_spmv_dot", A1.h, a, c_, a, 40, 41)
        want_y, want_s = solo.download(b_, ng), solo.scalars(40, 2)
        for transport, (y, sc) in got.items():
            same = np.array_equal(y, want_y)
            rel = float(np.max(np.abs(sc - want_s) / np.abs(want_s)))
            print(f"unstructured {ng} rows, {world} ranks, {'peer memory' if transport else 'NCCL'}: A(A(Ax)) "
                  f"{'bit-identical to' if same else 'DIFFERS from'} the single-GPU result; fused dots rel. diff {rel:.1e}", flush=True)
            ok &= same and rel <= 1e-13
dist.barrier()
if rank == 0:
    with capi.Context(local) as solo:
        for (method, pre), r in results.items():
            s = host.solve(solo, method, pre, matrix_name=name, want_x=False, max_iters=300)
            k = min(s.history.size, r.history.size)
            err = float(np.max(np.abs(s.history[:k] - r.history[:k])) / s.history[0])
            same = s.iter_count == r.iter_count and np.array_equal(s.history, r.history)
            print(f"{name} -{method} -p {pre}: its {r.iter_count} vs single-GPU {s.iter_count}, "
                  f"max |dr|/r0 = {err:.2e}: " + ("WHOLE HISTORY BIT-IDENTICAL" if same else "histories differ"), flush=True)
            if peer and 8 % world == 0:
                ok &= same
            else:
                kk = cmp_len(method, s.history.size, r.history.size)
                e2 = float(np.max(np.abs(s.history[:kk] - r.history[:kk])) / s.history[0])
                ok &= e2 <= 1e-10 and abs(r.iter_count - s.iter_count) <= max(2, (0.15 if method == "bi" else 0.05) * s.iter_count)
    print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
ctx.close()
dist.barrier()
dist.destroy_process_group()
