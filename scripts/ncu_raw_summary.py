#!/usr/bin/env python
"""Per-launch table from `ncu -i rep --page raw --csv`: duration, DRAM bytes / %, L2 %, tensor %, issue-slot %."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
def g(r, k):
    return r[hdr.index(k)] if k in hdr else ''
def f(x):
    try: return float(x)
    except ValueError: return float('nan')
print(f"{'id':>3} {'kernel':16s} {'grid':>4} {'us':>8} {'dramR_MB':>9} {'dramW_MB':>9} {'dram%':>6} {'l2%':>6} {'tensor%':>7} {'issue%':>6} {'smem%':>6} {'regs':>4}")
for r in rows[2:]:
    name = g(r, "Kernel Name").split("(")[0].split("::")[-1][:16]
    print(f"{g(r,'ID'):>3} {name:16s} {g(r,'launch__grid_size'):>4} {f(g(r,'gpu__time_duration.sum')):8.1f} {f(g(r,'dram__bytes_read.sum')):9.1f} {f(g(r,'dram__bytes_write.sum')):9.1f} "
          f"{f(g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} {f(g(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
          f"{f(g(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):7.1f} {f(g(r,'sm__inst_issued.avg.pct_of_peak_sustained_active')):6.1f} "
          f"{f(g(r,'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed')):6.1f} {g(r,'launch__registers_per_thread'):>4}")
print('units:', units[hdr.index('dram__bytes_read.sum')], units[hdr.index('gpu__time_duration.sum')])
