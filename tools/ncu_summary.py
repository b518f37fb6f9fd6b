"""Key metrics of one `ncu --set full` capture as a markdown table.  usage: python tools/ncu_summary.py raw.csv  (raw.csv = `ncu -i X.ncu-rep --page raw --csv`)"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(f"kernel: {r[hdr.index('Kernel Name')][:80]}")
    print("| metric | value |\n|---|---|")
    for k in KEYS:
        if k in hdr:
            print(f"| {k} | {r[hdr.index(k)]} {units[hdr.index(k)]} |")
