"""Summary of an `ncu --set full` raw-page CSV holding several kernels (ncu -i rep --page raw --csv): last launch of every
(kernel, grid) pair with the metrics DESIGN.md quotes.  usage: python scripts/summarize_others.py raw.csv > profiles/rNN_others.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__icc_request_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
KEYS = [k for k in KEYS if k in idx]
seen = collections.OrderedDict()
for r in rows[2:]:
    seen.setdefault((re.sub(r"\(.*", "", r[idx["Kernel Name"]]), r[idx["launch__grid_size"]]), []).append(r)
print("# " + " ".join(sys.argv[2:]))
for (name, g), rs in seen.items():
    r = rs[-1]
    print(f"\n{name}  grid {g}  ({len(rs)} launches captured; last one shown)")
    for k in KEYS:
        print(f"   {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
