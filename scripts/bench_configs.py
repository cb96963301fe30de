"""Device-resident throughput of the other BASELINE.json configurations (supplementary to bench.py, whose
headline line is configs[1]): configs[2] mixed AM / SAM / FM, configs[3] SSB + spectral NR + spectrum FFT,
configs[4] RX + SSB TX, plus wide SSB.  Inputs: a few distinct synthetic channels per kind, tiled.
usage: python scripts/bench_configs.py [--channels N] [--blocks T] [--steps K]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uhsdr_b200 import synth
from uhsdr_b200.config import (DEMOD_AM, DEMOD_FM, DEMOD_LSB, DEMOD_SAM, DSP_NR_ENABLE, SAM_SIDEBAND_USB, default_cfg)
from uhsdr_b200.engine import Engine
from uhsdr_b200.partition import partition_by_kind

PLANS = {
    "ssb_narrow": [("usb35", default_cfg()), ("lsb38", default_cfg(dmod_mode=DEMOD_LSB, filter_path=38))],
    "ssb_wide": [("usb48", default_cfg(filter_path=48)), ("usb55", default_cfg(filter_path=55))],
    "mixed_am_sam_fm": [("am70", default_cfg(dmod_mode=DEMOD_AM, filter_path=70)),
                        ("sam72", default_cfg(dmod_mode=DEMOD_SAM, filter_path=72)),
                        ("sam72u", default_cfg(dmod_mode=DEMOD_SAM, filter_path=72, sam_sideband=SAM_SIDEBAND_USB)),
                        ("fm2", default_cfg(dmod_mode=DEMOD_FM, filter_path=2))],
    "ssb_nr_spectrum": [("usb35nr", default_cfg(dsp_active=DSP_NR_ENABLE, spectrum_enable=1))],
}


def run(plan, nch, nb, steps, tx=False, spectrum=False):
    kinds = PLANS[plan]
    dev = torch.device("cuda", 0)
    per = nch // len(kinds)
    distinct = 8
    rows, cfg_of = [], []
    for k, (name, cfg) in enumerate(kinds):
        base = np.stack([synth.rx_iq(cfg, c, nb * 32, seed=11) for c in range(distinct)])
        rows.append(np.tile(base, ((per + distinct - 1) // distinct, 1, 1))[:per])
        cfg_of += [cfg] * per
    iq = torch.from_numpy(np.concatenate(rows)).to(dev)
    n = iq.shape[0]
    out = torch.empty_like(iq)
    with Engine(n) as eng:
        for k, (name, cfg) in enumerate(kinds):            # channels sorted by kind
            eng.configure(cfg, first=k * per, count=per)
        ext = torch.cuda.ExternalStream(eng.stream, device=dev)
        mic = iqtx = None
        if tx:
            mic = torch.from_numpy(np.tile(np.stack([synth.tx_mic(c, nb * 32) for c in range(distinct)]), ((n + distinct - 1) // distinct, 1, 1))[:n]).to(dev)
            iqtx = torch.empty_like(mic)
        mags = torch.empty((n, 512), dtype=torch.float32, device=dev) if spectrum else None

        def step():
            eng.rx_device(iq, out, nb)
            if tx:
                eng.tx_device(mic, iqtx, nb)
            if spectrum:                                   # one 512-point spectrum FFT per 512 input samples and channel
                for _ in range(max(1, nb * 32 // 512)):
                    eng._check(eng._lib.uhsdr_get_spectrum_device(eng._h, 0, n, mags.data_ptr()))
        with torch.cuda.stream(ext):
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(steps):
                step()
            e1.record(ext)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    units = n * nb * 32 * (2 if tx else 1)
    print(json.dumps({"plan": plan + ("+tx" if tx else ""), "channels": n, "blocks": nb, "ms_per_step": ms,
                      "channel_samples_per_s": units / (ms * 1e-3), "note": "device-resident, CUDA events, 1 GPU"}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=400)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    run("ssb_narrow", 4096, a.blocks, a.steps)
    run("ssb_wide", 4096, a.blocks, a.steps)
    run("mixed_am_sam_fm", 16384, a.blocks, a.steps)
    run("ssb_nr_spectrum", 4096, a.blocks, a.steps, spectrum=True)
    run("ssb_narrow", 8192, a.blocks, a.steps, tx=True)
