"""Summarise an `ncu --set full --import-source on` report of one kernel into a text file for profiles/.

usage: python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_name.txt ["note ..."]

Sections: headline metrics (duration, DRAM bytes, pipe/issue utilisation, registers, shared-memory
conflicts), then the SASS split at BAR.SYNC into warp roles with instruction counts, stall-sample
shares and the opcode mix -- the evidence DESIGN.md quotes.
"""
import csv
import io
import re
import subprocess
import sys

import pandas as pd

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__inst_executed_pipe_lsu.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def op(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s.strip())
    return s.split()[0].split(".")[0] if s else "?"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines.append(f"# ncu --set full summary of {rep.split('/')[-1]}  ({len(rows) - 2} profiled launch(es); last one shown)")
    if note:
        lines.append(f"# {note}")
    vals = rows[-1]
    name_i = hdr.index("Kernel Name")
    lines.append(f"kernel: {vals[name_i]}")
    for i, h in enumerate(hdr):
        if h in KEYS:
            lines.append(f"{h:72s} {units[i]:16s} {vals[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sl = src.splitlines()
    start = next(i for i, l in enumerate(sl) if l.startswith('"Address"'))
    df = pd.read_csv(io.StringIO("\n".join(sl[start:])))
    df["addr"] = df["Address"].apply(lambda x: int(str(x), 16))
    df = df.sort_values("addr").drop_duplicates("addr").reset_index(drop=True)
    df["op"] = df["Source"].apply(op)
    tot = max(1, df["# Samples"].sum())
    stalls = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
    lines.append("")
    lines.append("total warp-instructions executed: %.1f M; stall samples %d" % (df["Instructions Executed"].sum() / 1e6, tot))
    lines.append("whole-kernel stall shares (%): " + ", ".join(f"{k[6:]}={v:.1f}" for k, v in (df[stalls].sum().sort_values(ascending=False)[:8] / tot * 100).items()))
    lines.append("")
    lines.append("SASS split at BAR.SYNC (consecutive barriers delimit one warp role's loop body):")
    bars = df.index[df["Source"].str.contains("BAR.SYNC")].tolist()
    prev = 0
    for b in bars + [len(df) - 1]:
        w = df.iloc[prev:b + 1]
        n = w["Instructions Executed"].sum()
        if n > 0:
            top = w[stalls].sum().sort_values(ascending=False)[:5]
            h = w.groupby("op")["Instructions Executed"].sum().sort_values(ascending=False)
            lines.append(f"  sass {prev:5d}-{b:5d}: {n / 1e6:8.1f} M warp-inst, {w['# Samples'].sum() / tot * 100:5.1f}% of samples; stalls "
                         + ", ".join(f"{k[6:]}={v / tot * 100:.1f}" for k, v in top.items()))
            lines.append("      ops (M): " + ", ".join(f"{k}={v / 1e6:.1f}" for k, v in h.items() if v / n > 0.01))
        prev = b + 1
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
