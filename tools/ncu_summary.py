"""Key metrics + stall breakdown from an .ncu-rep (reads `ncu -i rep --page raw --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
names = [r[hdr.index("Kernel Name")][:60] for r in data]
for n, nm in enumerate(names):
    print(f"[{n}] {nm}")
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:72s} {units[i]:10s} " + "  ".join(f"{r[i]:>14s}" for r in data))
print("-- stalls per issue-active")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        vals = [float(r[i]) for r in data]
        if max(vals) >= 0.05:
            print(f"{h[34:-23]:30s} " + "  ".join(f"{v:8.3f}" for v in vals))
