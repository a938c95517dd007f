"""Summarise one kernel of an .ncu-rep (run where ncu is installed; no GPU needed):
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "sm__cycles_elapsed.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep}: first captured launch (ncu --set full --clock-control none --import-source on)")
for k in want:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k:90s} {units[i]:16s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [x for x in rows[2:] if len(x) >= len(hdr)]
tot = sum(int(x[idx["# Samples"]]) for x in data) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"\n# source page: top stall sites ({tot} samples, {len(data)} SASS instructions)")
for x in sorted(data, key=lambda x: -int(x[idx["# Samples"]]))[:14]:
    s = sorted(((h, int(x[idx[h]])) for h in stalls if int(x[idx[h]]) > 0), key=lambda kv: -kv[1])[:2]
    print(f"{int(x[idx['# Samples']]) * 100 / tot:5.1f}%  {x[idx['Source']].strip()[:60]:60s} {s}")
agg = {}
for x in data:
    for h in stalls:
        agg[h] = agg.get(h, 0) + int(x[idx[h]])
print("\n# stall reasons, all samples:", sorted(agg.items(), key=lambda kv: -kv[1])[:7])
