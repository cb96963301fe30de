"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel counts, total
device time and share.  usage: python scripts/summarize_launches.py gpurun_out/launches.csv profiles/out.txt ["note"]"""
import io
import sys

import pandas as pd

src, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
lines = open(src).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
df = pd.read_csv(io.StringIO("\n".join(lines[start:])))
df = df[df["Metric Name"] == "gpu__time_duration.sum"].copy()
df["ns"] = df["Metric Value"].astype(str).str.replace(",", "").astype(float)
df["kernel"] = df["Kernel Name"].str.replace(r"\(.*", "", regex=True)
g = df.groupby("kernel")["ns"].agg(["count", "sum", "mean"]).sort_values("sum", ascending=False)
tot = g["sum"].sum()
res = [f"# launch list {src.split('/')[-1]}: {len(df)} launches, {tot / 1e6:.3f} ms device time (cold-cache, serialised under ncu: compare shares)"]
if note:
    res.append(f"# {note}")
res.append(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'mean us':>10s} {'share %':>8s}")
for k, r in g.iterrows():
    res.append(f"{k[:60]:60s} {int(r['count']):8d} {r['sum'] / 1e6:10.3f} {r['mean'] / 1e3:10.1f} {100 * r['sum'] / tot:8.2f}")
open(out, "w").write("\n".join(res) + "\n")
print("\n".join(res))
