"""Summarise an ncu --set full report of the fused kernel by kernel role: the SASS between two
consecutive BAR.SYNC instructions belongs to one warp role (FIR | LAT | AGC | EQ | POST)."""
import io
import subprocess
import sys

import pandas as pd

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(__import__("csv").reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"]
for i, h in enumerate(hdr):
    if h in keys:
        print(f"{h:70s} {rows[1][i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
df = pd.read_csv(io.StringIO("\n".join(lines[1:])))
df["addr"] = df["Address"].apply(lambda x: int(str(x), 16))
df = df.sort_values("addr").reset_index(drop=True)
tot = df["# Samples"].sum()
bars = df.index[df["Source"].str.contains("BAR.SYNC")].tolist()
stalls = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
prev = 0
print("total warp-instructions executed: %.1f M, samples %d" % (df["Instructions Executed"].sum() / 1e6, tot))
for b in bars + [len(df) - 1]:
    w = df.iloc[prev:b + 1]
    top = w[stalls].sum().sort_values(ascending=False)[:6]
    print(f"sass {prev:5d}-{b:5d}: samples {w['# Samples'].sum() / tot * 100:5.1f}%  inst {w['Instructions Executed'].sum() / 1e6:8.1f} M  "
          + ", ".join(f"{k[6:]}={v / tot * 100:.1f}" for k, v in top.items()))
    prev = b + 1
print({k[6:]: round(v, 1) for k, v in (df[stalls].sum().sort_values(ascending=False)[:8] / tot * 100).items()})
