"""A/B timing of two builds of the CUDA library in ONE process on ONE box (clocks differ between boxes and calls, so
only same-call ratios mean anything): the narrow-SSB headline workload through uhsdr_rx_process_device, alternating
between the libraries.
usage: python scripts/ab_time.py libA.so libB.so [--channels 4096] [--blocks 1500] [--rounds 3] [--plan narrow|wide|mixed]"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uhsdr_b200.config import (DEMOD_AM, DEMOD_FM, DEMOD_LSB, DEMOD_SAM, DSP_NR_ENABLE, SAM_SIDEBAND_USB, ChanCfg, default_cfg)   # noqa: E402
from uhsdr_b200.tables import DEFAULT_BLOB   # noqa: E402

def _rep(c):
    return [c, c, c, c]


PLANS = {
    "narrow": [default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)],
    "wide": [default_cfg(filter_path=48), default_cfg(filter_path=48), default_cfg(filter_path=55), default_cfg(filter_path=55)],
    "mixed": [default_cfg(dmod_mode=DEMOD_AM, filter_path=70), default_cfg(dmod_mode=DEMOD_SAM, filter_path=72),
              default_cfg(dmod_mode=DEMOD_SAM, filter_path=72, sam_sideband=SAM_SIDEBAND_USB), default_cfg(dmod_mode=DEMOD_FM, filter_path=2)],
    "am": _rep(default_cfg(dmod_mode=DEMOD_AM, filter_path=70)), "sam": _rep(default_cfg(dmod_mode=DEMOD_SAM, filter_path=72)),
    "samu": _rep(default_cfg(dmod_mode=DEMOD_SAM, filter_path=72, sam_sideband=SAM_SIDEBAND_USB)), "fm": _rep(default_cfg(dmod_mode=DEMOD_FM, filter_path=2)),
    "w48": _rep(default_cfg(filter_path=48)), "w55": _rep(default_cfg(filter_path=55)),
    "nr": _rep(default_cfg(dsp_active=DSP_NR_ENABLE)),
}


class Lib:
    def __init__(self, path, nch, cfgs):
        # "lib.so::NAME=VALUE,NAME2=VALUE2": environment switches read at engine creation (UHSDR_B200_NO_SERIAL2=1 ...)
        path, _, envs = path.partition("::")
        saved = {}
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        self.L = L = ctypes.CDLL(os.path.abspath(path))
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.uhsdr_engine_create.argtypes = [ctypes.POINTER(vp), ci, ci, vp, ctypes.c_size_t]
        L.uhsdr_configure_channels_strided.argtypes = [vp, ci, ci, ci, ctypes.POINTER(ChanCfg), ci]
        L.uhsdr_rx_process_device.argtypes = [vp, vp, vp, vp, ci, vp]
        L.uhsdr_tx_process_device.argtypes = [vp, vp, vp, vp, ci, vp]
        L.uhsdr_engine_stream.restype = vp
        L.uhsdr_engine_stream.argtypes = [vp]
        L.uhsdr_engine_sync.argtypes = [vp]
        L.uhsdr_engine_destroy.argtypes = [vp]
        blob = open(DEFAULT_BLOB, "rb").read()
        buf = ctypes.create_string_buffer(blob, len(blob))
        self.h = vp()
        assert L.uhsdr_engine_create(ctypes.byref(self.h), nch, 0, buf, len(blob)) == 0
        k = len(cfgs)
        for i, c in enumerate(cfgs):
            if k == 2:      # alternating plan (BASELINE configs[1])
                cnt = (nch - i + k - 1) // k
                assert L.uhsdr_configure_channels_strided(self.h, i, cnt, k, ctypes.byref(c), 1) == 0
            else:           # channels sorted by kind
                lo, hi = i * nch // k, (i + 1) * nch // k
                assert L.uhsdr_configure_channels_strided(self.h, lo, hi - lo, 1, ctypes.byref(c), 1) == 0
        self.stream = torch.cuda.ExternalStream(L.uhsdr_engine_stream(self.h), device=torch.device("cuda", 0))
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def run(self, iq, out, nb, reps, tx=False):
        L = self.L
        if tx:      # the modulator alone (TxProcessor_Run): iq = microphone words, out = I/Q words
            call = lambda: L.uhsdr_tx_process_device(self.h, iq.data_ptr(), out.data_ptr(), None, nb, None)
            with torch.cuda.stream(self.stream):
                assert call() == 0
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
                for _ in range(reps):
                    assert call() == 0
                e1.record(self.stream)
                torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        with torch.cuda.stream(self.stream):
            assert L.uhsdr_rx_process_device(self.h, iq.data_ptr(), out.data_ptr(), None, nb, None) == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(reps):
                assert L.uhsdr_rx_process_device(self.h, iq.data_ptr(), out.data_ptr(), None, nb, None) == 0
            e1.record(self.stream)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--channels", type=int, default=4096)
    ap.add_argument("--blocks", type=int, default=1500)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--plan", default="narrow")
    ap.add_argument("--tx", action="store_true", help="time uhsdr_tx_process_device instead of the receiver")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    iq = (torch.randn((a.channels, a.blocks * 32, 2), device=dev, generator=g) * 3000.0 * 65536.0).to(torch.int32)
    if a.tx:
        iq[:, :, 1] = iq[:, :, 0]
    out = torch.empty_like(iq)
    libs = [Lib(p, a.channels, PLANS[a.plan]) for p in a.libs]
    ms = [[] for _ in libs]
    for _ in range(a.rounds):
        for i, lb in enumerate(libs):
            ms[i].append(lb.run(iq, out, a.blocks, a.reps, a.tx))
    best = [min(m) for m in ms]
    print(json.dumps({"plan": a.plan, "channels": a.channels, "blocks": a.blocks,
                      "ms": {os.path.basename(p) + f"#{i}": [round(x, 4) for x in m] for i, (p, m) in enumerate(zip(a.libs, ms))},
                      "chsamp_per_s": {os.path.basename(p) + f"#{i}": a.channels * a.blocks * 32 / (b * 1e-3) for i, (p, b) in enumerate(zip(a.libs, best))},
                      "speedup_vs_first": [round(best[0] / b, 4) for b in best]}), flush=True)


if __name__ == "__main__":
    main()
