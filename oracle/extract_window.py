"""ORACLE / BUILD-TIME TOOL.  Extracts the function-local constant `von_Hann_1024`
(reference: mchf-eclipse/drivers/ui/lcd/ui_spectrum.c:362) into a raw float32 file so that
dump_tables can append it to the table blob.  Reads the reference source, writes only numbers."""
import re
import sys

import numpy as np

src = open(sys.argv[1], encoding="latin-1").read()
m = re.search(r"von_Hann_1024\s*\[1024\]\s*=\s*\{(.*?)\};", src, re.S)
vals = np.array([float(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()], dtype=np.float32)
assert vals.size == 1024, vals.size
vals.tofile(sys.argv[2])
