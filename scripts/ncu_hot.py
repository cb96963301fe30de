"""Top stall sites of an ncu --set full --import-source on report: SASS index, samples, dominant stall reasons, source line.
usage: python scripts/ncu_hot.py rep.ncu-rep [lo hi] [N]"""
import io, subprocess, sys
import pandas as pd
rep = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
N = int(sys.argv[4]) if len(sys.argv) > 4 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sl = src.splitlines()
start = next(i for i, l in enumerate(sl) if l.startswith('"Address"'))
df = pd.read_csv(io.StringIO("\n".join(sl[start:])))
df["addr"] = df["Address"].apply(lambda x: int(str(x), 16))
df = df.sort_values("addr").drop_duplicates("addr").reset_index(drop=True)
stalls = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
w = df.iloc[lo:hi + 1]
tot = df["# Samples"].sum()
print("region samples: %.1f%% of kernel" % (w["# Samples"].sum() / tot * 100))
for idx, r in w.sort_values("# Samples", ascending=False).head(N).iterrows():
    top = r[stalls].astype(float).sort_values(ascending=False)[:3]
    print(f"{idx:5d} {r['# Samples'] / tot * 100:5.2f}% exec={int(r['Instructions Executed']):9d}  " + ",".join(f"{k[6:]}={int(v)}" for k, v in top.items() if v > 0) + "   " + str(r["Source"])[:90])
