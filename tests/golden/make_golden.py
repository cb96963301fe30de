"""Generates the committed golden vectors tests/golden/rx_golden.npz (and tx/spectrum ones) by running
the REFERENCE's own object code (oracle/_ref/libuhsdr_ref.so, built from /root/reference by
oracle/Makefile) on seeded synthetic inputs.  Needs /root/reference; run from the repo root:

    make -C oracle ref && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import FM_TONE_CASES, NB_CASES, NB_IMPULSES, NR_CASES, RX_CASES, SPECTRUM_CASES, TX_CASES  # noqa: E402
from oracle.refchain import RefChannel  # noqa: E402
from uhsdr_b200 import synth  # noqa: E402
from uhsdr_b200.config import default_cfg  # noqa: E402


def main():
    out = {}
    for label, kw, nblocks in RX_CASES + NR_CASES + NB_CASES:
        cfg = default_cfg(**kw)
        iq = synth.rx_iq(cfg, 5, nblocks * 32, seed=1234)
        if (label, kw, nblocks) in NB_CASES:
            iq = synth.add_impulses(iq, 1234, NB_IMPULSES)
        with RefChannel(cfg) as r:
            audio, audio_f = r.rx(iq)
            st = r.status()
        out[f"{label}/iq"] = iq
        out[f"{label}/audio_l"] = audio[:, 0].copy()
        out[f"{label}/audio_f"] = audio_f
        out[f"{label}/status"] = np.array([st.adc_clip, st.agc_action, st.fm_squelched, st.sam_carrier_freq_offset], dtype=np.int32)
        assert np.array_equal(audio[:, 0], audio[:, 1])
    for label, kw, nblocks, tone in FM_TONE_CASES:
        cfg = default_cfg(**kw)
        iq = synth.rx_fm_subtone_iq(cfg, 5, nblocks * 32, 100.0, 300.0 if tone else 0.0, seed=1234)
        with RefChannel(cfg) as r:
            audio, audio_f = r.rx(iq)
        out[f"{label}/iq"] = iq
        out[f"{label}/audio_l"] = audio[:, 0].copy()
        out[f"{label}/audio_f"] = audio_f
    # mute + reconfigure sequence on one channel
    cfg_a, cfg_b = default_cfg(), default_cfg(filter_path=44, bass_gain=0)
    iq = synth.rx_iq(cfg_a, 9, 160 * 32, seed=77)
    mute = np.zeros(160, dtype=np.uint8)
    mute[40:48] = 1
    with RefChannel(cfg_a) as r:
        a1, f1 = r.rx(iq[: 80 * 32], mute[:80])
        r.reconfigure(cfg_b)
        a2, f2 = r.rx(iq[80 * 32:], mute[80:])
    out["seq_mute_reconf/iq"] = iq
    out["seq_mute_reconf/mute"] = mute
    out["seq_mute_reconf/audio_l"] = np.concatenate([a1[:, 0], a2[:, 0]])
    out["seq_mute_reconf/audio_f"] = np.concatenate([f1, f2])
    # spectrum-display FFT (UiSpectrum_RedrawSpectrum states 0-2) after 37 and after 100 blocks
    for label, kw in SPECTRUM_CASES:
        cfg = default_cfg(**kw)
        z = 1 << cfg.spectrum_magnify                # the zoom FFT decimates: as many ring entries as without it
        iq = synth.rx_iq(cfg, 4, 100 * z * 32, seed=55)
        with RefChannel(cfg) as r:
            r.rx(iq[: 37 * z * 32])
            m1 = r.spectrum()
            r.rx(iq[37 * z * 32:])
            m2 = r.spectrum()
        out[f"{label}/iq"] = iq
        out[f"{label}/mags37"] = m1
        out[f"{label}/mags100"] = m2
    # SSB transmit chain
    for label, kw, nblocks in TX_CASES:
        cfg = default_cfg(**kw)
        mic = synth.tx_mic(6, nblocks * 32, seed=99)
        mute = np.zeros(nblocks, dtype=np.uint8)
        mute[nblocks // 2: nblocks // 2 + 5] = 1
        with RefChannel(cfg) as r:
            iqw, iqf = r.tx(mic, mute)
            st = r.status()
        out[f"{label}/mic"] = mic[:, 0].copy()
        out[f"{label}/mute"] = mute
        out[f"{label}/iq"] = iqw
        out[f"{label}/iq_f"] = iqf
        out[f"{label}/status"] = np.array([st.tx_alc_val, st.tx_peak_audio], dtype=np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rx_golden.npz"), **out)
    print("wrote rx_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
