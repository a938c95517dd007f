"""Residual histories of the UNMODIFIED compiled reference (oracle/_ref/libbis_ref.so) at the
BASELINE.json sizes -- histories only, a few kB per case.

Run from the repo root, one case per process (they take minutes to tens of minutes on one core):

    python tests/golden/make_golden_large.py hpcg128_sgs
    python tests/golden/make_golden_large.py all        # every case, sequentially

Each case writes tests/golden/large_<case>.npz with
    history   collected residual norms of the 1-thread run (the golden)
    meta      [iter_count, converged, restarts]
    final     [final true residual, stopping criterion]
    noise     [max_k |r_k(8 threads) - r_k(1 thread)| / r0, min its, max its]  (reference self-noise)
The -m gpu parity tests (tests/test_baseline_sizes_gpu.py) compare the device histories with these.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import matgen, refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# case -> (matrix builder, method, preconditioner)
CASES = {
    # BASELINE configs[0]
    "hpcg128_cg": (lambda: matgen.hpcg(128), "cg", "none"),
    # configs[1]
    "hpcg128_j": (lambda: matgen.hpcg(128), "j", "none"),
    "hpcg128_sgs": (lambda: matgen.hpcg(128), "sgs", "none"),
    # configs[2]
    "hpcg256_cg_sgs": (lambda: matgen.hpcg(256), "cg", "sgs"),
    # configs[3] (Anderson content is this repo's restatement, SCAMAC absent: the SAME CRS goes to both sides)
    "anderson_gm_j": (lambda: matgen.anderson(100, 100, 50, ranpot=5.0, t=1.0, seed=1, periodic=False), "gm", "j"),
    "anderson_gm_ilu0": (lambda: matgen.anderson(100, 100, 50, ranpot=5.0, t=1.0, seed=1, periodic=False), "gm", "ilu0"),
    # configs[4] proxies: the reference cannot hold HPCG-512 (32-bit nnz)
    "hpcg256_bi_j": (lambda: matgen.hpcg(256), "bi", "j"),
    "hpcg256_cg_j": (lambda: matgen.hpcg(256), "cg", "j"),
    "hpcg128_bi_j": (lambda: matgen.hpcg(128), "bi", "j"),
    "hpcg128_cg_sgs": (lambda: matgen.hpcg(128), "cg", "sgs"),
}


def run(case: str) -> None:
    build, method, pre = CASES[case]
    t0 = time.time()
    rp, col, val = build()
    print(f"{case}: matrix {rp.size - 1} rows, {col.size} nnz ({time.time() - t0:.1f} s)", flush=True)
    r = refshim.solve(rp, col, val, method, pre, threads=1)
    print(f"  1 thread: its={r.iter_count} conv={int(r.converged)} restarts={r.restarts} r0={r.history[0]:.16e} "
          f"r1={r.history[1]:.16e} final={r.final_true_residual:.6e} ({time.time() - t0:.1f} s)", flush=True)
    q = refshim.solve(rp, col, val, method, pre, threads=8)
    k = min(q.history.size, r.history.size)
    with np.errstate(invalid="ignore"):
        noise = float(np.nanmax(np.abs(q.history[:k] - r.history[:k])) / r.history[0])
    print(f"  8 threads: its={q.iter_count} self-noise={noise:.2e} ({time.time() - t0:.1f} s)", flush=True)
    np.savez_compressed(
        os.path.join(OUT, f"large_{case}.npz"),
        history=r.history,
        meta=np.array([r.iter_count, int(r.converged), r.restarts], np.int64),
        final=np.array([r.final_true_residual, r.stopping_criteria]),
        noise=np.array([noise, min(r.iter_count, q.iter_count), max(r.iter_count, q.iter_count)]),
        history8=q.history,
    )


if __name__ == "__main__":
    assert refshim.available(), "build oracle/_ref first: make -C oracle ref"
    which = sys.argv[1:] or ["all"]
    if which == ["all"]:
        which = list(CASES)
    for c in which:
        run(c)
