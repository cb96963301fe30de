import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from uhsdr_b200 import synth
from uhsdr_b200.config import default_cfg
from uhsdr_b200.engine import Engine
cfg = default_cfg(filter_path=48)
nch, nb = 4096, 100
base = np.stack([synth.rx_iq(cfg, c, nb * 32, seed=11) for c in range(8)])
iq = torch.from_numpy(np.tile(base, (nch // 8, 1, 1))).cuda()
out = torch.empty_like(iq)
with Engine(nch) as eng:
    eng.configure(cfg)
    for _ in range(3):
        eng.rx_device(iq, out, nb)
    eng.sync()
print("ok")
