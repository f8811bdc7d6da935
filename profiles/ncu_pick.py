"""Selected metrics of an `ncu -i report.ncu-rep --page raw --csv` export as a small CSV (metric, unit, value).
usage: ncu -i x.ncu-rep --page raw --csv > raw.csv; python profiles/ncu_pick.py raw.csv > profiles/<name>_ncu.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct")
w = csv.writer(sys.stdout)
w.writerow(("metric", "unit", "value"))
for h, u, v in zip(hdr, units, vals):
    if h in keep or ("issue_stalled" in h and h.endswith("_per_issue_active.ratio")):
        w.writerow((h, u, v))
