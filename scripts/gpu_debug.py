import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import refchain
from uhsdr_b200 import synth
from uhsdr_b200.config import *
from uhsdr_b200.engine import Engine

def run(label, cfg, nblocks=64, split=False):
    n = nblocks * 32
    iq = synth.rx_iq(cfg, 0, n)[None]
    with Engine(1, exact=True) as e:
        e.configure(cfg)
        if split:
            audio = np.concatenate([e.rx(iq[:, :n // 2]), e.rx(iq[:, n // 2:])], axis=1)
        else:
            audio = e.rx(iq)
    with refchain.RefChannel(cfg) as r:
        ref_a, ref_f = r.rx(iq[0])
    got = audio[0, :, 0] >> 16
    want = ref_a[:, 0] >> 16
    bad = np.flatnonzero(got != want)
    print(label, "split" if split else "single", "first mismatch", bad[:5], "of", bad.size)
    if bad.size:
        i = bad[0]
        lo = max(0, i - 4)
        print("  got ", got[lo:lo + 16])
        print("  want", want[lo:lo + 16])

run("p55 agc off", default_cfg(filter_path=55, agc_mode=5))
run("p35 agc off", default_cfg(agc_mode=5))
run("p35", default_cfg())
run("p35", default_cfg(), split=True)
