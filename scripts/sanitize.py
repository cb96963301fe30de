"""Small invocation of every kernel of the library through the host-buffer C ABI (no torch), for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|initcheck python scripts/sanitize.py [--exact] [--only NAME]
Covers ragged CTAs (30 and 61 channels), short calls (4 blocks) and calls that are not a multiple of 4 or 16 blocks, the
tensor-core kernel, the CUDA-core fused kernel, the split general path (front2 + serial, NR, notch), the one-kernel general
path, TX (SSB / AM / FM, with and without a mute array), spectrum FFT + display states, the twin-peaks kernel and
reconfiguration.  Prints a checksum per case so a run can be compared with a run without the tool."""
import argparse
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uhsdr_b200 import synth                                                  # noqa: E402
from uhsdr_b200.config import (DEMOD_AM, DEMOD_FM, DEMOD_LSB, DEMOD_SAM, DSP_NB_ENABLE, DSP_NOTCH_ENABLE, DSP_NR_ENABLE,   # noqa: E402
                               default_cfg, default_spectrum_display_cfg)
from uhsdr_b200.engine import Engine                                          # noqa: E402


def iq_for(cfgs, nb):
    return np.concatenate([synth.counter_block(np, [synth.kind_of(c)], [i], 0, nb * 32) for i, c in enumerate(cfgs)])


def case_rx(name, cfgs, nbs, exact, mute=False, reconf=None):
    n = len(cfgs)
    with Engine(n, exact=exact) as eng:
        for i, c in enumerate(cfgs):
            eng.configure(c, first=i, count=1)
        crc = 0
        for k, nb in enumerate(nbs):
            iq = iq_for(cfgs, nb)
            m = None
            if mute:
                m = np.zeros((n, nb), dtype=np.uint8); m[:, nb // 2:] = 1
            out = eng.rx(iq, m)
            crc = zlib.crc32(out.tobytes(), crc)
            if reconf is not None and k == 0:
                eng.configure(reconf, first=0, count=n, reset=False)
        st = eng.status()
    print(f"{name:28s} channels={n:3d} calls={nbs} crc={crc:08x} tw={st[0].twinpeaks_state}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    usb, lsb = default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)
    cases = {
        "tc_ragged_30": lambda: case_rx("tc_ragged_30", [usb, lsb] * 15, [8, 4, 12], a.exact),
        "tc_61_mute": lambda: case_rx("tc_61_mute", [usb] * 61, [16, 8], a.exact, mute=True),
        "narrow_not_mult_of_4": lambda: case_rx("narrow_not_mult_of_4", [usb, lsb] * 5, [7, 9], a.exact),
        "narrow_aa_treble": lambda: case_rx("narrow_aa_treble", [default_cfg(filter_path=44, treble_gain=3)] * 9, [8, 8], a.exact),
        "wide_split": lambda: case_rx("wide_split", [default_cfg(filter_path=48), default_cfg(filter_path=55)] * 17, [20, 5, 64], a.exact),
        "am_sam_fm": lambda: case_rx("am_sam_fm", [default_cfg(dmod_mode=DEMOD_AM, filter_path=70), default_cfg(dmod_mode=DEMOD_SAM, filter_path=72),
                                                   default_cfg(dmod_mode=DEMOD_FM, filter_path=2), default_cfg(dmod_mode=DEMOD_SAM, filter_path=84)] * 9, [18, 70], a.exact, mute=True),
        "nr_nb_notch": lambda: case_rx("nr_nb_notch", [default_cfg(dsp_active=DSP_NR_ENABLE), default_cfg(dsp_active=DSP_NR_ENABLE | DSP_NB_ENABLE, nb_setting=6),
                                                       default_cfg(dsp_active=DSP_NOTCH_ENABLE)] * 4, [40, 72], a.exact),
        "reconfigure": lambda: case_rx("reconfigure", [usb] * 6, [12, 12], a.exact, reconf=default_cfg(filter_path=44, bass_gain=0)),
    }

    def tx_case():
        cfgs = [usb, lsb, default_cfg(dmod_mode=DEMOD_AM, filter_path=70), default_cfg(dmod_mode=DEMOD_FM, filter_path=2)] * 8 + [usb]
        n = len(cfgs)
        with Engine(n, exact=a.exact) as eng:
            for i, c in enumerate(cfgs):
                eng.configure(c, first=i, count=1)
            crc = 0
            for nb, with_mute in ((20, False), (6, True), (70, False)):
                mic = np.concatenate([synth.counter_block(np, [synth.KIND_MIC], [i], 0, nb * 32) for i in range(n)])
                m = None
                if with_mute:
                    m = np.zeros((n, nb), dtype=np.uint8); m[:, 2:4] = 1
                crc = zlib.crc32(eng.tx(mic, m).tobytes(), crc)
        print(f"{'tx_ssb_am_fm':28s} channels={n:3d} crc={crc:08x}", flush=True)

    def spectrum_case():
        cfgs = [default_cfg(spectrum_enable=1), default_cfg(spectrum_enable=1, spectrum_magnify=2, filter_path=48)] * 3 + [usb]
        n = len(cfgs)
        with Engine(n, exact=a.exact) as eng:
            for i, c in enumerate(cfgs):
                eng.configure(c, first=i, count=1)
            eng.rx(iq_for(cfgs, 40))
            mags = eng.spectrum()
            disp, lvl, avg = eng.spectrum_display(default_spectrum_display_cfg())
            eng.twinpeaks_rearm()
            eng.rx(iq_for(cfgs, 8))
        print(f"{'spectrum_display':28s} channels={n:3d} crc={zlib.crc32(mags.tobytes() + disp.tobytes() + lvl.tobytes()):08x}", flush=True)

    cases["tx_ssb_am_fm"] = tx_case
    cases["spectrum_display"] = spectrum_case
    for name, fn in cases.items():
        if a.only is None or a.only == name:
            fn()
    print("sanitize.py done", flush=True)


if __name__ == "__main__":
    main()
