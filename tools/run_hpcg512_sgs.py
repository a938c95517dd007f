"""HPCG-512 -cg -p sgs on ONE GPU (possible since the factors give up their natural-order CRS):
   python tools/run_hpcg512_sgs.py [n] [iters]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi, host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
with capi.Context(0) as ctx:
    t0 = time.time()
    sess = host.BenchSession(ctx, f"HPCG-{n}", "cg", "sgs")
    sess.prepare(3)
    t_setup = time.time() - t0
    free, total = torch.cuda.mem_get_info(0)
    r = sess.run(K)
    h = sess.history(3 + K + 1)
    print(f"HPCG-{n} -cg -p sgs on one GPU: set-up + 3 iterations {t_setup:.1f} s, device memory in use {(total - free) / 2**30:.1f} GiB "
          f"of {total / 2**30:.1f}; {r['device_ms'] / K:.3f} ms/iter over {K} iterations; residuals {h[0]:.6e} -> {h[-1]:.6e} "
          f"(peak {h.max():.6e} at iteration {int(h.argmax())}: preconditioned CG's 2-norm residual rises first, as in the reference's HPCG-256 run)")
    sess.close()
