"""Warp-stall breakdown and per-opcode executed-instruction mix of one kernel in an ncu report:
   python tools/ncu_stalls.py gpurun_out/<name>.ncu-rep ["<note>"]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
print(f"# {rep}  {note}")
want = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.per_cycle_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
stall = sorted(((float(v), h) for h, v in zip(hdr, vals)
                if "average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)
print("warp cycles per issued instruction, by stall reason:")
for v, h in stall[:10]:
    print(f"  {v:7.3f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter()
for r in rows[2:]:
    if len(r) <= ix["Instructions Executed"]:
        continue
    ex = int(r[ix["Instructions Executed"]] or 0)
    parts = r[ix["Source"]].split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    mix[op.split(".")[0]] += ex
tot = sum(mix.values())
print(f"executed warp instructions by opcode (total {tot}):")
for op, n in mix.most_common(24):
    print(f"  {op:10s} {n:10d}  {100.0 * n / tot:5.1f} %")
