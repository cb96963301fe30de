"""Development check (run on the GPU box): CUDA engine vs the oracle for a list of configurations."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import refchain
from oracle.port import PortChannel
from uhsdr_b200 import synth
from uhsdr_b200.config import *
from uhsdr_b200.engine import Engine


def oracle_rx(cfg, iq):
    if refchain.available():
        with refchain.RefChannel(cfg) as r:
            return r.rx(iq)
    with PortChannel(cfg) as p:
        return p.rx(iq)


def check(label, cfg, nblocks=600, nch=3, exact=False):
    n = nblocks * 32
    iq = np.stack([synth.rx_iq(cfg, c, n) for c in range(nch)])
    with Engine(nch, exact=exact) as e:
        e.configure(cfg)
        t0 = time.time()
        # two calls to exercise state carry-over
        h = (nblocks // 2) * 32
        a1 = e.rx(iq[:, :h]); a2 = e.rx(iq[:, h:])
        audio = np.concatenate([a1, a2], axis=1)
        dt = time.time() - t0
    worst = 0.0
    exact_all = True
    for c in range(nch):
        ref_a, ref_f = oracle_rx(cfg, iq[c])
        got = audio[c, :, 0].astype(np.float64) / 65536.0
        want = ref_a[:, 0].astype(np.float64) / 65536.0
        err = got - want
        rms = np.sqrt(np.mean(want ** 2)) + 1e-30
        snr = 10 * np.log10(np.mean(want ** 2) / (np.mean(err ** 2) + 1e-30))
        rel = np.max(np.abs(err)) / (np.max(np.abs(want)) + 1e-30)
        worst = max(worst, rel)
        exact_all &= np.array_equal(audio[c], ref_a)
        if c == 0:
            print(f"{label:28s} {'EXACT' if exact else 'fast '} ch0: rms {rms:9.2f} maxrel {rel:.2e} snr {snr:6.1f} dB bit-exact={np.array_equal(audio[c], ref_a)} ({dt*1e3:.0f} ms)")
    return worst, exact_all


if __name__ == "__main__":
    cfgs = [
        ("USB p35", default_cfg()),
        ("LSB p38", default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)),
        ("USB p44 (aa)", default_cfg(filter_path=44)),
        ("USB p48 (hil first)", default_cfg(filter_path=48)),
        ("USB p55 (M2)", default_cfg(filter_path=55)),
        ("USB p65 (M2 aa)", default_cfg(filter_path=65)),
        ("CW p8", default_cfg(dmod_mode=DEMOD_CW, filter_path=8)),
        ("AM p70", default_cfg(dmod_mode=DEMOD_AM, filter_path=70)),
        ("SAM p72 both", default_cfg(dmod_mode=DEMOD_SAM, filter_path=72)),
        ("SAM p72 usb", default_cfg(dmod_mode=DEMOD_SAM, filter_path=72, sam_sideband=SAM_SIDEBAND_USB)),
        ("SAM p84 lsb", default_cfg(dmod_mode=DEMOD_SAM, filter_path=84, sam_sideband=SAM_SIDEBAND_LSB)),
        ("FM p2", default_cfg(dmod_mode=DEMOD_FM, filter_path=2)),
        ("USB +6k", default_cfg(iq_freq_mode=FREQ_IQ_CONV_P6KHZ)),
        ("USB +12k", default_cfg(iq_freq_mode=FREQ_IQ_CONV_P12KHZ)),
        ("USB manual iq", default_cfg(iq_auto_correction=0, rx_adj_gain_i=1.01, rx_adj_gain_q=0.99, iq_phase_balance_rx=-0.01)),
        ("USB agc off", default_cfg(agc_mode=5)),
        ("USB agc fast hang", default_cfg(agc_mode=4, agc_hang_enable=1)),
        ("USB notch/peak/eq", default_cfg(dsp_active=DSP_MNOTCH_ENABLE | DSP_MPEAK_ENABLE, treble_gain=3, bass_gain=-4)),
    ]
    for label, cfg in cfgs:
        for exact in (True, False):
            nb = 1200 if cfg.dmod_mode == DEMOD_FM else 600
            check(label, cfg, nblocks=nb, exact=exact)
