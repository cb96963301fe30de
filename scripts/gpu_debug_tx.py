import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch
from uhsdr_b200 import synth
from uhsdr_b200.config import *
from uhsdr_b200.engine import Engine
g = np.load("tests/golden/rx_golden.npz")
cfg = default_cfg()
nb = 160
mic = synth.tx_mic(6, nb * 32, seed=99)
mute = g["tx_usb/mute"]
want = g["tx_usb/iq"]
for nch in (1, 2, 3, 5):
    for exact in (True, False):
        with Engine(nch, exact=exact) as e:
            e.configure(cfg)
            got = e.tx(np.stack([mic] * nch), np.stack([mute] * nch))
        for c in range(nch):
            bad = np.flatnonzero(np.any(got[c] != want, axis=1))
            print("nch", nch, "exact", exact, "ch", c, "mismatch", bad[:3], bad.size)
