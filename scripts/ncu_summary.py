"""Summarise an .ncu-rep: key metrics + top stall instructions. usage: ncu_summary.py file.ncu-rep [topN]"""
import csv, subprocess, sys, io
f = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_pipe_fp64.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {units[i]:12s} {vals[i][:100]}")
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.3:
        print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:22s} {float(vals[i]):.2f}")
src = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
data = [(int(r[isamp]) if r[isamp].isdigit() else 0, i, r[isrc].strip(), r[iex]) for i, r in enumerate(rows[2:])]
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
for s, i, text, ex in sorted(data, reverse=True)[:topn]:
    print(f"{s:7d} {100*s/tot:5.1f}%  #{i:4d} ex={ex:>9s}  {text[:80]}")
