"""Per-role view of an ncu --set full --import-source on report of the fused kernels: the SASS between two BAR.SYNC is one
warp role's loop body; samples taken while a warp waits at the barrier land on the instruction after it and are
attributed back to the role that waits.  Prints busy share, stall mix, executed (hot) code size per role, and the
instruction-cache hit rate.   usage: python scripts/ncu_roles.py rep.ncu-rep"""
import csv, io, subprocess, sys
import pandas as pd
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
for k in ("gpu__time_duration.sum", "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread"):
    if k in h: print(f"{k:70s} {v[h.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sl = src.splitlines()
start = next(i for i, l in enumerate(sl) if l.startswith('"Address"'))
df = pd.read_csv(io.StringIO("\n".join(sl[start:])))
df["addr"] = df["Address"].apply(lambda x: int(str(x), 16))
df = df.sort_values("addr").drop_duplicates("addr").reset_index(drop=True)
df.to_pickle("/tmp/ncu_last.pkl")
tot = df["# Samples"].sum()
stalls = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
bars = df.index[df["Source"].str.contains("BAR.SYNC")].tolist()
prev = 0
carry = 0.0
for b in bars + [len(df) - 1]:
    w = df.iloc[prev:b + 1]
    bar_here = w["stall_barrier"].sum() if "stall_barrier" in w else 0
    own = w["# Samples"].sum() - bar_here
    n = w["Instructions Executed"].sum()
    hot = (w["Instructions Executed"] > 5000).sum()
    if prev > 0:
        print(f"   (barrier wait of the previous role: {bar_here / tot * 100:.1f}%)")
    if n > 0:
        top = (w[[c for c in stalls if c != "stall_barrier"]].sum().sort_values(ascending=False)[:6])
        print(f"sass {prev:5d}-{b:5d}: {n / 1e6:6.1f} M inst, hot code {hot * 16 / 1024:5.1f} KB, busy samples {own / tot * 100:5.1f}%: "
              + ", ".join(f"{k[6:]}={x / max(own, 1) * 100:.0f}%" for k, x in top.items()))
    prev = b + 1
