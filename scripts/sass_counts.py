"""Static SASS evidence: instruction counts per kernel of the shipping library (cuobjdump -sass), for the mnemonics that prove
which hardware path a kernel uses (tcgen05 = UTCHMMA / UTCBAR / LDTM, mbarrier = SYNCS, bulk async copy = UBLKCP, 256-bit global
accesses = LDG/STG.E.ENL2.256).  usage: python scripts/sass_counts.py [lib.so] > profiles/rNN_sass_counts.txt"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "uhsdr_b200/csrc/libuhsdr_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOM", "SYNCS", "UBLKCP", "LDG.E.ENL2.256", "STG.E.ENL2.256", "LDG.E.128", "STG.E.128",
       "LDS.128", "STS.128", "FFMA", "F2FP", "MUFU", "BAR", "SHFL"]
print(f"# cuobjdump -sass {lib}: static instruction counts per kernel (mnemonic prefix match)")
print("%-34s %7s " % ("kernel", "instrs") + " ".join(pat))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", f)
    m = re.search(r"uhsdr\d+([a-z0-9_]+?_kernel)", name)
    short = (m.group(1) if m else name)[:30] + ("<1>" if "ILb1E" in name else "<0>" if "ILb0E" in name else "")
    c = [sum(1 for i in ins if re.search(r"(^|\s)" + re.escape(p), i)) for p in pat]
    print("%-34s %7d " % (short, len(ins)) + " ".join("%*d" % (len(p), x) for p, x in zip(pat, c)))
