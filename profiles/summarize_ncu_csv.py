#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` (a file) into the counters DESIGN.md / bench.py cite, one block per launch."""
import csv
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "smsp__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
print(f"# {sys.argv[1]} ({len(rows) - 2} launch(es)); ncu --set full --clock-control none; cold-cache, serialised")
for r in rows[2:]:
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:85s} {r[i]} {units[i]}")
    print("-" * 100)
