"""Per-role timing of rx_ssb_tc_kernel (tools only): runs the narrow-SSB workload on libuhsdr_b200_prof.so
(`make -C uhsdr_b200/csrc prof`), which carries clock64() timers around the work section of every warp role of
CTA 0 and a knock-out mask that switches roles off (results wrong, timing only).
usage: python scripts/tc_role_times.py [--channels 4096] [--blocks 400] [--masks 0,1,2,...]"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uhsdr_b200.engine as E                                      # noqa: E402

E.LIB_FAST = os.path.join(ROOT, "uhsdr_b200", "csrc", "libuhsdr_b200_prof.so")
from uhsdr_b200.config import DEMOD_LSB, default_cfg               # noqa: E402

ROLES = ["fe0", "fe1", "fe2", "fe3", "fe4", "fe5", "fe6", "postA", "epi0", "epi1", "epi2", "epi3", "bq", "postB", "lat", "agc", "mma"]
KNOCKS = {1: "front end", 2: "MMAs", 4: "dec epilogue", 8: "hil epilogue", 16: "gain law", 32: "lattice", 64: "AGC", 128: "biquads", 256: "output"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=4096)
    ap.add_argument("--blocks", type=int, default=400)
    ap.add_argument("--masks", default="0,1,2,4,8,16,32,64,128,256,3,15,31,479,447,255")
    ap.add_argument("--pauses", default="-1", help="MMA issue pause overrides (cycles) to sweep, -1 = built-in")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    iq = (torch.randn((a.channels, a.blocks * 32, 2), device=dev, generator=g) * 3000.0 * 65536.0).to(torch.int32)
    out = torch.empty_like(iq)
    eng = E.Engine(a.channels)
    lib = eng._lib
    lib.uhsdr_debug_tc_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
    eng.configure(default_cfg(), first=0, stride=2)
    eng.configure(default_cfg(dmod_mode=DEMOD_LSB, filter_path=38), first=1, stride=2)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    buf = (ctypes.c_ulonglong * 128)()
    runs = [(int(m), int(pz)) for pz in a.pauses.split(",") for m in a.masks.split(",")]
    for mask, pz in runs:
        assert lib.uhsdr_debug_tc_pause(pz) == 0
        assert lib.uhsdr_debug_tc_prof(None, mask) == 0
        with torch.cuda.stream(ext):
            for _ in range(2):
                eng.rx_device(iq, out, a.blocks)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(3):
                eng.rx_device(iq, out, a.blocks)
            e1.record(ext)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        assert lib.uhsdr_debug_tc_prof(buf, mask) == 0
        v = np.array(list(buf), dtype=np.float64).reshape(32, 4)
        per = {ROLES[w]: round(v[w, 0] / max(v[w, 1], 1)) for w in range(len(ROLES))}
        step = round(v[0, 2] / max(v[0, 1] - 1, 1))
        off = [KNOCKS[b] for b in KNOCKS if mask & b]
        print(json.dumps({"knocked_out": off, "mma_pause": pz, "ms": round(ms, 4), "chsamp_per_s": a.channels * a.blocks * 32 / (ms * 1e-3),
                          "cycles_per_step": step, "work_cycles_per_step": per}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
