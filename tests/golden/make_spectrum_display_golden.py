"""Generates tests/golden/spectrum_display_golden.npz by running the REFERENCE's own spectrum state machine
(UiSpectrum_RedrawSpectrum states 0-4, ui_spectrum.c compiled into oracle/_ref/libuhsdr_ref.so behind
oracle/ref_spectrum_harness.c) on seeded synthetic I/Q.  Needs /root/reference; run from the repo root:

    make -C oracle ref && python tests/golden/make_spectrum_display_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import SPECDISP_BLOCKS, SPECDISP_CASES, SPECDISP_REDRAWS  # noqa: E402
from oracle.refchain import RefChannel  # noqa: E402
from uhsdr_b200 import synth  # noqa: E402
from uhsdr_b200.config import default_cfg, default_spectrum_display_cfg  # noqa: E402


def main():
    out = {}
    for label, kw, dkw in SPECDISP_CASES:
        cfg, dc = default_cfg(**kw), default_spectrum_display_cfg(**dkw)
        z = 1 << cfg.spectrum_magnify
        n = SPECDISP_BLOCKS * z * 32
        iq = synth.rx_iq(cfg, 4, SPECDISP_REDRAWS * n, seed=91)
        out[f"{label}/iq"] = iq
        with RefChannel(cfg) as r:
            r.spectrum_display_init(dc)
            for k in range(SPECDISP_REDRAWS):
                r.rx(iq[k * n:(k + 1) * n])
                mags, avg, disp, lvl = r.spectrum_display()
                out[f"{label}/mags{k}"], out[f"{label}/avg{k}"], out[f"{label}/disp{k}"], out[f"{label}/lvl{k}"] = mags, avg, disp, lvl
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "spectrum_display_golden.npz"), **out)
    print("wrote spectrum_display_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
