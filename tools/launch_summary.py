"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
   python tools/launch_summary.py gpurun_out/launches_bench512.csv "<command that was profiled>" > profiles/rNN_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    for line in f:
        if line.startswith('"'):
            rows.append(line)
r = list(csv.reader(rows))
hdr = r[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
acc = OrderedDict()
for row in r[1:]:
    if row[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", row[ki])
    name = re.sub(r"<unnamed>::", "", name)
    m = re.search(r"(bis_\w+)::\[lambda", row[ki])
    if m:
        name = name.split("<")[0] + f"<{m.group(1)}>"
    t = float(row[vi].replace(",", "")) / 1e6
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in acc.values())
print(f"ncu launch list of: {sys.argv[2] if len(sys.argv) > 2 else '?'} (cold-cache, serialised: compare shares)")
for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} launches {t:10.2f} ms {100 * t / tot:5.1f}%  avg {t / n:9.3f} ms  {name[:110]}")
g = lambda pat: next(((n, t) for k, (n, t) in acc.items() if pat in k), (0, 0.0))
sp, cu, cd = g("EpiDot"), g("bis_cg_update"), g("bis_cg_direction")
if sp[0] and cu[0] and cd[0]:
    a, b, c = sp[1] / sp[0], cu[1] / cu[0], cd[1] / cd[0]
    print(f"\nper CG iteration (avg launch durations): spmv_dot {a:.2f} ms + cg_update {b:.2f} ms + cg_direction {c:.2f} ms "
          f"= {a + b + c:.2f} ms; SpMV share {a / (a + b + c):.3f}")
