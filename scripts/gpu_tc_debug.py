"""Debug driver for the tensor-core fused kernel: N channels x nb blocks, compares against the
CUDA-core fused kernel (UHSDR_B200_NO_TC=1 engine in the same process is not possible: the switch
is read at engine creation, so two engines are created with the variable toggled)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uhsdr_b200 import synth
from uhsdr_b200.config import DEMOD_LSB, default_cfg
from uhsdr_b200.engine import Engine

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg_u, cfg_l = default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)
base = np.stack([synth.rx_iq(cfg_u if c % 2 == 0 else cfg_l, c, nb * 32, seed=8) for c in range(8)])
iq = np.tile(base, ((nch + 7) // 8, 1, 1))[:nch]
dev = torch.device("cuda", 0)
d_iq = torch.from_numpy(iq).to(dev)
res = {}
for tc in (1, 0):
    os.environ["UHSDR_B200_NO_TC"] = "0" if tc else "1"
    with Engine(nch) as eng:
        eng.configure(cfg_u, first=0, stride=2)
        if nch > 1:
            eng.configure(cfg_l, first=1, stride=2)
        out = torch.empty_like(d_iq)
        f = torch.empty(d_iq.shape[:2], dtype=torch.float32, device=dev)
        eng.rx_device(d_iq, out, nb, audio_f_dev=f)
        eng.sync()
        res[tc] = (out.cpu().numpy(), f.cpu().numpy())
        print("tc" if tc else "fused", "ok", flush=True)
a, b = res[1][1].astype(np.float64), res[0][1].astype(np.float64)
err = a - b
print("max abs err", np.abs(err).max(), "peak", np.abs(b).max(), "snr dB", 10 * np.log10((b ** 2).mean() / max((err ** 2).mean(), 1e-300)))
