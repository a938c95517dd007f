"""Per-row timeline of one forward solve on HPCG-n (debug aid):
   BIS_TRSV_DEBUG_FILE=gpurun_out/trsv_trace.bin python tools/trsv_trace.py n"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basic_iterative_solvers_b200 import capi, host  # noqa: E402

n = int(sys.argv[1])
path = os.environ["BIS_TRSV_DEBUG_FILE"]
rp, col, val = host.matrix(f"HPCG-{n}")
f = host.factor(rp, col, val, "sgs")
N = len(rp) - 1
with capi.Context(0) as ctx:
    L = ctx.upload_triangular(f.l_rp, f.l_col, f.l_val, upper=False)
    D, b, x = ctx.upload(f.A_D), ctx.upload(np.ones(N)), ctx.alloc(N)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.set_option("trsv_debug", 1)
    ctx.call("bis_sptrsv", L.h, x, D, b)
    ctx.sync()
raw = np.fromfile(path, dtype=np.uint64).reshape(N, 4)
ts = raw[:, :3].astype(np.int64)
spins = (raw[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
first_round = (raw[:, 3] >> np.uint64(32)).astype(np.int64)
t0 = ts[:, 2].min()
ts -= t0
np.save(path + ".npy", ts[:, 2].astype(np.int32))
np.save(path + ".all.npy", np.concatenate([ts, spins[:, None], first_round[:, None]], axis=1).astype(np.int32))
print("poll rounds that found an operand missing: median %d mean %.1f p90 %d ; first round took (ns) median %d p90 %d" %
      (np.median(spins), spins.mean(), np.percentile(spins, 90), np.median(first_round[spins > 0]), np.percentile(first_round[spins > 0], 90)))
w = ts[:, 1] - ts[:, 0]
print("time in the final wait loop (ns): median %d ; per round: median %.0f" % (np.median(w), np.median(w[spins > 0] / (spins[spins > 0] + 1))))
os.remove(path)
idx = np.arange(N)
lvl = idx % n + 2 * ((idx // n) % n) + 4 * (idx // (n * n))      # level(x,y,z) = x + 2y + 4z
nl = lvl.max() + 1
first_pub = np.full(nl, 1 << 62)
last_pub = np.zeros(nl, np.int64)
np.minimum.at(first_pub, lvl, ts[:, 2])
np.maximum.at(last_pub, lvl, ts[:, 2])
print(f"HPCG-{n}: {nl} levels, total {(ts[:,2].max())/1e3:.1f} us")
d_last = np.diff(last_pub)
print("per-level advance of the LAST publication (ns): median %.0f mean %.0f p90 %.0f" % (np.median(d_last), d_last.mean(), np.percentile(d_last, 90)))
print("row: crit observed -> operands loaded (ns): median %.0f ; operands -> published: median %.0f" %
      (np.median(ts[:, 1] - ts[:, 0]), np.median(ts[:, 2] - ts[:, 1])))
# hop: time from the publication of a row's critical dependency (x-1, same line) to its own publication
has = idx % n > 0
hop = ts[has, 2] - ts[idx[has] - 1, 2]
print("hop (x-1 published -> row published) ns: median %.0f mean %.0f p10 %.0f p90 %.0f" % (np.median(hop), hop.mean(), np.percentile(hop, 10), np.percentile(hop, 90)))
obs = ts[has, 0] - ts[idx[has] - 1, 2]
print("x-1 published -> critical dep observed by the row (ns): median %.0f p10 %.0f p90 %.0f" % (np.median(obs), np.percentile(obs, 10), np.percentile(obs, 90)))
for l in (10, 100, nl // 2, nl - 100):
    print(f"level {l}: rows {np.sum(lvl == l)}, first pub {first_pub[l]/1e3:.2f} us, last pub {last_pub[l]/1e3:.2f} us, spread {(last_pub[l]-first_pub[l])/1e3:.2f} us")
