#!/usr/bin/env python
"""bench.py -- headline benchmark: channel-samples/s of the full SSB RX chain (BASELINE.json).

A "step" is one pass of the hot path (uhsdr_rx_process_device: AudioDriver_RxProcessor for every
channel) over one batch of synthetic I/Q: `--channels` channels per GPU (default 4096, BASELINE.json
configs[1]: alternating USB / LSB on FilterPathInfo[35] / [38]) x `--blocks` consecutive 32-sample
blocks (default 1500 = 1 s of signal per channel).  One channel-sample = one 48 ksps complex input
sample fully processed to its audio output sample.

  value      whole-job channel-samples/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the host-buffer C-ABI call (uhsdr_rx_process): pinned host I/Q in,
             pinned host audio out, H2D + D2H copies inside the timed region
  roofline   algorithmic bytes (16 B per channel-sample) / launch duration vs the measured HBM peak,
             plus the FP32-pipe view (346 FLOP per channel-sample)
  cpu_baseline / --impl reference
             the reference's own C chain (oracle/_ref, else the oracle port) on the host cores

Multi-GPU: one process per GPU (torchrun), channels partitioned by rank, no collective on the data
path ("scaling": "weak"); torch.distributed is used only for the barrier and the max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = 346.0     # narrow SSB chain, MAC = 2 FLOP
BYTES_PER_SAMPLE = 16.0     # 8 B IqSample_t in + 8 B AudioSample_t out
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def measured_traffic():
    """DRAM bytes per channel-sample of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except OSError:
        return None


def channel_cfg(ch: int):
    from uhsdr_b200.config import DEMOD_LSB, DEMOD_USB, default_cfg
    return default_cfg(dmod_mode=DEMOD_USB, filter_path=35) if ch % 2 == 0 else default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)


def plan_groups(name: str):
    """Channel plans of BASELINE.json's configs: list of (label, cfg, share) -- channels are sorted by kind
    (8e) and every kind gets `share` of them; "ssb_narrow" alternates instead (configs[1] as written)."""
    from uhsdr_b200.config import (DEMOD_AM, DEMOD_FM, DEMOD_LSB, DEMOD_SAM, DSP_NR_ENABLE, SAM_SIDEBAND_USB, default_cfg)
    if name in ("ssb_narrow", "rx_tx"):
        return [("usb_p35", default_cfg(), 1), ("lsb_p38", default_cfg(dmod_mode=DEMOD_LSB, filter_path=38), 1)]
    if name == "ssb_wide":
        return [("usb_p48", default_cfg(filter_path=48), 1), ("usb_p55", default_cfg(filter_path=55), 1)]
    if name == "mixed_am_sam_fm":
        return [("am_p70", default_cfg(dmod_mode=DEMOD_AM, filter_path=70), 2), ("sam_p72_both", default_cfg(dmod_mode=DEMOD_SAM, filter_path=72), 1),
                ("sam_p72_usb", default_cfg(dmod_mode=DEMOD_SAM, filter_path=72, sam_sideband=SAM_SIDEBAND_USB), 1),
                ("fm_p2", default_cfg(dmod_mode=DEMOD_FM, filter_path=2), 2)]
    if name == "ssb_nr_spectrum":
        return [("usb_p35_nr_spec", default_cfg(dsp_active=DSP_NR_ENABLE, spectrum_enable=1), 1)]
    raise KeyError(name)


# direct-form FLOP per channel-sample (MAC = 2), SURVEY.md 8a / 8d
PLAN_FLOP = {"ssb_narrow": 346.0, "ssb_wide": 440.0, "mixed_am_sam_fm": 310.0, "ssb_nr_spectrum": 346.0 + 65.0, "rx_tx": (346.0 + 910.0) / 2}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(idx, nblocks, reps, barrier, q):
    from oracle import refchain
    from oracle.port import PortChannel
    from uhsdr_b200 import synth
    cfg = channel_cfg(idx)
    iq = synth.counter_block(np, [synth.kind_of(cfg)], [idx], 0, nblocks * 32)[0]
    ch = refchain.RefChannel(cfg) if refchain.available() else PortChannel(cfg)
    ch.rx(iq[: 32 * 64])    # warm the caches / page in the library
    times = []
    for _ in range(reps):
        barrier.wait()
        t0 = time.perf_counter()
        ch.rx(iq)
        times.append(time.perf_counter() - t0)
        barrier.wait()
    q.put((idx, times))


def run_cpu_chain(cores: int, nblocks: int, reps: int):
    """`cores` independent channels, one process each, `reps` synchronised passes of nblocks blocks.
    Returns the list of per-pass wall times (max over workers)."""
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores)
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(i, nblocks, reps, barrier, q)) for i in range(cores)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    per_pass = [max(r[1][k] for r in res) for k in range(reps)]
    return per_pass


def cpu_kind():
    from oracle import refchain
    return "reference" if refchain.available() else "port"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nblocks = 15000          # 10 s of signal per channel and step
    per_pass = run_cpu_chain(cores, nblocks, args.warmup + args.steps)[args.warmup:]
    total = sum(per_pass)
    samples = cores * nblocks * 32 * args.steps
    value = samples / total
    line = {
        "metric": "channel-samples/s, full SSB RX chain", "value": value, "unit": "channel-samples/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SSB RX chain, alternating USB path 35 / LSB path 38 (BASELINE.json configs[1] channel mix), "
                               f"{cores} channels x {nblocks} blocks per step on the host CPU", "channels": cores, "blocks_per_step": nblocks},
        "cpu_baseline": {"value": value, "unit": "channel-samples/s", "cores": cores, "kind": cpu_kind(),
                         "sample": f"{cores} channels x {nblocks} blocks x {args.steps} steps, one process per channel"},
        "e2e": {"value": value, "unit": "channel-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms (B200_PROFILING.md recipe).  Started
    before the warm-up (nvidia-smi needs a few hundred ms to come up); only the samples whose
    timestamps fall inside [t_begin, t_end] -- the timed region plus, when that is shorter than
    MIN_LOAD_S, an untimed continuation of the very same launches -- are summarised."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    MIN_LOAD_S = 1.5

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t_begin: float, t_end: float):
        import datetime
        if self.p is not None:
            time.sleep(0.15)
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons, total = [], [], [], set(), 0
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(parts[2]), float(parts[3])
            except ValueError:
                continue
            total += 1
            if ts < t_begin - 0.05 or ts > t_end + 0.05:
                continue
            sm.append(clk); smax.append(cmax)
            try:
                power.append(float(parts[4]))
            except ValueError:
                pass
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[6:10]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_total": total}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None,
                "window": "timed region + untimed continuation of the same launches (>= %.1f s under load)" % self.MIN_LOAD_S}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
DTYPE_LABEL = "f32 (I/O, state and accumulation f32; operands of the decimator / Hilbert FIR GEMMs split into 2 bf16 terms)"


def local_plan(name: str, nch: int, ch0: int):
    """Per-rank channel plan: list of (cfg, label) per local channel, channels of one kind contiguous (except the alternating
    narrow-SSB plan of configs[1]), and the generator kind of each channel."""
    from uhsdr_b200 import synth
    groups = plan_groups(name)
    cfgs, labels = [], []
    if name in ("ssb_narrow", "rx_tx"):
        for c in range(nch):
            lab, cfg, _ = groups[(ch0 + c) % 2]
            cfgs.append(cfg); labels.append(lab)
    else:
        tot = sum(g[2] for g in groups)
        done = 0
        for gi, (lab, cfg, share) in enumerate(groups):
            cnt = nch - done if gi == len(groups) - 1 else (nch * share) // tot
            cfgs += [cfg] * cnt; labels += [lab] * cnt
            done += cnt
    kinds = [synth.kind_of(c) for c in cfgs]
    return cfgs, labels, kinds


def configure_plan(eng, name, cfgs, ch0):
    if name in ("ssb_narrow", "rx_tx") and len(cfgs) >= 2:
        first_usb = 0 if ch0 % 2 == 0 else 1
        eng.configure(cfgs[first_usb], first=first_usb, stride=2)
        eng.configure(cfgs[1 - first_usb], first=1 - first_usb, stride=2)
        return
    c = 0
    while c < len(cfgs):
        e = c
        while e < len(cfgs) and cfgs[e] is cfgs[c]:
            e += 1
        eng.configure(cfgs[c], first=c, count=e - c)
        c = e


def gen_device(torch, kinds, ch_global, ns, dev, mic=False, chunk=128):
    """Counter-based integer generator (uhsdr_b200/synth.py counter_block), evaluated on the device: the same code on numpy
    regenerates any channel / time slice bit for bit on the host."""
    from uhsdr_b200 import synth
    out = torch.empty((len(kinds), ns, 2), dtype=torch.int32, device=dev)
    for c0 in range(0, len(kinds), chunk):
        k = [synth.KIND_MIC] * min(chunk, len(kinds) - c0) if mic else kinds[c0:c0 + chunk]
        out[c0:c0 + len(k)] = synth.counter_block(torch, k, ch_global[c0:c0 + len(k)], 0, ns, device=dev)
    return out


def snr_db(got, ref):
    e = got.astype(np.float64) - ref.astype(np.float64)
    return float(10 * np.log10(max(np.mean(ref.astype(np.float64) ** 2), 1e-30) / max(np.mean(e ** 2), 1e-30)))


def parity_at_size(torch, eng, name, cfgs, labels, iq, mic, P, nsub, dev, tx, spectrum, ch0):
    """The first P blocks of the FULL batch from fresh state; a subset of channels that covers every kind is compared with
    the oracle (the compiled reference where oracle/_ref travelled, else the port) on the float audio before the int32
    formatting: max |err| / max |ref| (north_star: <= 1e-4) and SNR (>= 90 dB), plus the int32 words (<= 1 LSB of 16 bits)."""
    from oracle import refchain
    from oracle.port import PortChannel
    nch = len(cfgs)
    configure_plan(eng, name, cfgs, ch0)
    sub_iq = iq[:, : P * 32].contiguous()
    out = torch.empty_like(sub_iq)
    out_f = torch.empty((nch, P * 32), dtype=torch.float32, device=dev)
    eng.rx_device(sub_iq, out, P, audio_f_dev=out_f)
    if tx:
        sub_mic = mic[:, : P * 32].contiguous()
        txo = torch.empty_like(sub_mic)
        txo_f = torch.empty((nch, P * 32, 2), dtype=torch.float32, device=dev)
        eng.tx_device(sub_mic, txo, P, iq_f_dev=txo_f)
    eng.sync()
    # subset: evenly spaced, at least 8 of every kind
    idx = set(np.linspace(0, nch - 1, num=min(nsub, nch), dtype=np.int64).tolist())
    for lab in dict.fromkeys(labels):
        members = [i for i, l in enumerate(labels) if l == lab]
        idx.update(members[:: max(1, len(members) // 8)][:8])
    idx = sorted(idx)
    rel, snr, lsb, tx_rel, tx_snr, spec_rel = 0.0, 1e9, 0, 0.0, 1e9, 0.0
    per_kind = {}
    mags = eng.spectrum() if spectrum else None
    for c in idx:
        x = sub_iq[c].cpu().numpy()
        ch = refchain.RefChannel(cfgs[c]) if refchain.available() else PortChannel(cfgs[c])
        want, want_f = ch.rx(x)
        got_f = out_f[c].cpu().numpy()
        got_w = out[c, :, 0].cpu().numpy()
        r = float(np.max(np.abs(got_f - want_f)) / max(float(np.max(np.abs(want_f))), 1e-30))
        s = snr_db(got_f, want_f)
        rel, snr = max(rel, r), min(snr, s)
        lsb = max(lsb, int(np.max(np.abs((got_w.astype(np.int64) >> 16) - (want[:, 0].astype(np.int64) >> 16)))))
        k = per_kind.setdefault(labels[c], {"channels": 0, "max_rel_err": 0.0, "min_snr_db": 1e9})
        k["channels"] += 1; k["max_rel_err"] = max(k["max_rel_err"], r); k["min_snr_db"] = min(k["min_snr_db"], s)
        if spectrum:
            wm = ch.spectrum()
            spec_rel = max(spec_rel, float(np.max(np.abs(mags[c] - wm)) / np.max(wm)))
        if tx:
            wiq, wiq_f = ch.tx(sub_mic[c].cpu().numpy())
            g = txo_f[c].cpu().numpy()
            tx_rel = max(tx_rel, float(np.max(np.abs(g - wiq_f)) / max(float(np.max(np.abs(wiq_f))), 1e-30)))
            tx_snr = min(tx_snr, snr_db(g, wiq_f))
        ch.close()
    res = {"channels_compared": len(idx), "of": nch, "blocks": P, "oracle": cpu_kind(), "max_rel_err": rel, "min_snr_db": snr,
           "max_word_diff_lsb16": lsb, "per_kind": per_kind,
           "tolerance": "north_star: |err| <= 1e-4 relative (read as max|err| / max|ref| over the run) and >= 90 dB SNR, on the float audio",
           "within_tolerance": bool(rel <= 1e-4 and snr >= 90.0)}
    if tx:
        res["tx"] = {"max_rel_err": tx_rel, "min_snr_db": tx_snr}
        res["within_tolerance"] = bool(res["within_tolerance"] and tx_rel <= 1e-4 and tx_snr >= 90.0)
    if spectrum:
        res["spectrum_max_rel_err"] = spec_rel
    return res



def run_config(torch, dist, args, name, nch, T, rank, local_rank, world, dev, steps, warmup, tx=False, spectrum=False, keep=False):
    """One BASELINE config on this rank: engine + device-resident synthetic input, W warm-up and K timed steps (CUDA events on
    the engine's stream, barrier + synchronize on both sides, max over ranks), then the at-size parity check on rank 0."""
    from uhsdr_b200.engine import Engine
    from uhsdr_b200.partition import channel_range
    ns = T * 32
    ch0, ch1 = channel_range(rank, world, world * nch)
    cfgs, labels, kinds = local_plan(name, nch, ch0)
    ch_global = list(range(ch0, ch1))
    eng = Engine(nch, device=local_rank)
    configure_plan(eng, name, cfgs, ch0)
    iq = gen_device(torch, kinds, ch_global, ns, dev)
    audio = torch.empty_like(iq)
    mic = iqtx = mags = None
    if tx:
        mic = gen_device(torch, kinds, ch_global, ns, dev, mic=True)
        iqtx = torch.empty_like(mic)
    if spectrum:
        mags = torch.empty((nch, 512), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.rx_device(iq, audio, T)
        if tx:
            eng.tx_device(mic, iqtx, T)
        if spectrum:                       # one 512-point spectrum FFT per 512 input samples and channel
            for _ in range(max(1, ns // 512)):
                eng._check(eng._lib.uhsdr_get_spectrum_device(eng._h, 0, nch, mags.data_ptr()))

    with torch.cuda.stream(ext):
        for _ in range(warmup):
            step()
        # untimed spin-up of the same launches (bounded): a process that starts on an idle GPU reaches its sustained clock / power
        # state only after some tenths of a second under load; the W warm-up steps alone are a few milliseconds
        t_spin = time.time()
        for _ in range(400):
            if time.time() - t_spin >= args.spinup_s:
                break
            step()
            torch.cuda.synchronize()
        barrier()
        l0 = eng.launch_count
        t_begin = time.time()
        ev0.record(ext)
        for _ in range(steps):
            step()
        ev1.record(ext)
        barrier()
        launches = eng.launch_count - l0
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / steps
    units_per_gpu = nch * ns * (2 if tx else 1)
    res = {"name": name, "eng": eng, "iq": iq, "audio": audio, "ext": ext, "ms_per_step": ms_per_step, "launches": launches,
           "units_per_gpu": units_per_gpu, "value": float(world) * units_per_gpu / (ms_per_step * 1e-3), "t_begin": t_begin,
           "cfgs": cfgs, "labels": labels, "kinds": kinds, "ch0": ch0, "parity": None}
    if rank == 0 and args.parity_channels > 0:
        with torch.cuda.stream(ext):
            res["parity"] = parity_at_size(torch, eng, name, cfgs, labels, iq, mic, min(T, args.parity_blocks) // 4 * 4, args.parity_channels, dev, tx, spectrum, ch0)
            configure_plan(eng, name, cfgs, ch0)
    if not keep:
        eng.close()
        res["eng"] = None
        res["iq"] = res["audio"] = None
    return res


def roofline_of(res, peaks, which, plan):
    gbs = res["units_per_gpu"] * BYTES_PER_SAMPLE / (res["ms_per_step"] * 1e-3) / 1e9
    tf = res["units_per_gpu"] * PLAN_FLOP[plan] / (res["ms_per_step"] * 1e-3) / 1e12
    return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
            "peak_source": which, "algorithmic_bytes_per_step": res["units_per_gpu"] * BYTES_PER_SAMPLE,
            "fp32": {"achieved_tflops": tf, "peak_tflops_nominal": FP32_PEAK_TFLOPS_NOMINAL, "frac": tf / FP32_PEAK_TFLOPS_NOMINAL,
                     "flop_per_unit": PLAN_FLOP[plan]}}


def gpu_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the engine has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # host buffers of the e2e leg live on the NUMA node next to this rank's GPU: bind the process to the CPUs NVML names
    # for the device before anything is allocated (first touch places the pinned pages); without it 8 ranks share one node
    numa = None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa = sorted(os.sched_getaffinity(0))
        numa = f"{numa[0]}-{numa[-1]} ({len(numa)} cpus)"
    except Exception as exc:                                   # affinity is an optimisation, never a requirement
        numa = f"not set: {exc!r}"

    nch, T = args.channels, args.blocks
    ns = T * 32

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- headline: BASELINE.json configs[1] -------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    head = run_config(torch, dist, args, "ssb_narrow", nch, T, rank, local_rank, world, dev, args.steps, args.warmup, keep=True)
    eng, iq, audio, ext = head["eng"], head["iq"], head["audio"], head["ext"]
    with torch.cuda.stream(ext):
        # keep the same load running (untimed) until the clock sampler has seen MIN_LOAD_S of it
        t_load = time.time()
        while rank == 0 and time.time() - t_load < ClockSampler.MIN_LOAD_S:
            for _ in range(4):
                eng.rx_device(iq, audio, T)
            torch.cuda.synchronize()
        clocks = sampler.stop(head["t_begin"], time.time()) if rank == 0 else None
        barrier()
    ms_per_step, value, launches = head["ms_per_step"], head["value"], head["launches"]
    total_samples = float(world) * nch * ns

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h_iq = torch.empty((nch, ns, 2), dtype=torch.int32, pin_memory=True)
    h_audio = torch.empty((nch, ns, 2), dtype=torch.int32, pin_memory=True)
    h_iq.copy_(iq)
    torch.cuda.synchronize()
    lib, h = eng._lib, eng._h
    with torch.cuda.stream(ext):
        for _ in range(3):                                                               # warm-up (staging allocation, first touch)
            rc = lib.uhsdr_rx_process(h, h_iq.data_ptr(), h_audio.data_ptr(), T, None)
            assert rc == 0, rc
        barrier()
        ev0.record(ext)
        for _ in range(e2e_steps):
            rc = lib.uhsdr_rx_process(h, h_iq.data_ptr(), h_audio.data_ptr(), T, None)
            assert rc == 0, rc
        ev1.record(ext)
        barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if dist is not None:
        tms = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms.item())
    e2e_value = total_samples / (e2e_ms / e2e_steps * 1e-3)
    io_bytes = nch * ns * 8
    # the host-buffer path (sliced, three streams) must deliver what the device path delivers: compare a strided sample of rows
    e2e_match = None
    if rank == 0:
        with torch.cuda.stream(ext):
            configure_plan(eng, "ssb_narrow", head["cfgs"], head["ch0"])
            rc = lib.uhsdr_rx_process(h, h_iq.data_ptr(), h_audio.data_ptr(), T, None)
            assert rc == 0, rc
            configure_plan(eng, "ssb_narrow", head["cfgs"], head["ch0"])
            eng.rx_device(iq, audio, T)
            eng.sync()
            rows = torch.arange(0, nch, max(1, nch // 64), device=dev)
            e2e_match = bool(torch.equal(audio[rows].cpu(), h_audio[rows.cpu()]))
    del h_iq, h_audio
    eng.close()
    del iq, audio
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs (device-resident, same timing rules, at-size parity) -----------------------------
    others = []
    if not args.no_other_configs:
        osteps, owarm = max(1, min(args.steps, args.other_steps)), 3
        plans = [("ssb_wide", 4096, args.other_blocks, {}), ("mixed_am_sam_fm", 16384, args.other_blocks, {}),
                 ("ssb_nr_spectrum", 4096, args.other_blocks, {"spectrum": True}),
                 ("rx_tx", (65536 // world) if world > 1 else 8192, args.other_blocks, {"tx": True})]
        for pname, pch, pT, kw in plans:
            r = run_config(torch, dist, args, pname, pch, pT, rank, local_rank, world, dev, osteps, owarm, **kw)
            if rank == 0:
                peaks, which = measured_peaks()
                cfgdesc = {"ssb_wide": "wide SSB: Hilbert pair at 48 ksps, then the audio decimator (FilterPathInfo[48] / [55])",
                           "mixed_am_sam_fm": "BASELINE.json configs[2]: 1/3 AM path 70, 1/3 SAM path 72 (sideband BOTH and USB), 1/3 FM path 2, channels sorted by mode",
                           "ssb_nr_spectrum": "BASELINE.json configs[3]: SSB path 35 + spectral noise reduction + one 512-point spectrum FFT per 512 input samples and channel",
                           "rx_tx": "BASELINE.json configs[4]: narrow SSB RX chain + SSB TX modulator on every channel" +
                                    (f", 65536 channels sharded over {world} GPUs (contiguous ranges)" if world > 1 else ", one GPU's share at 8 GPUs (8192 channels)")}[pname]
                others.append({"name": pname, "workload": cfgdesc, "channels_per_gpu": pch, "channels_total": pch * world, "blocks_per_step": pT,
                               "unit": "channel-samples/s" + (" (RX + TX samples)" if kw.get("tx") else ""), "value": r["value"],
                               "ms_per_step": r["ms_per_step"], "steps": osteps, "warmup": owarm, "gpu_launches": r["launches"],
                               "roofline": roofline_of(r, peaks, which, pname), "parity": r["parity"]})
            torch.cuda.empty_cache()

    # ---- cpu baseline (rank 0, N == 1 only) -------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        reps, nb = 10, 15000
        one = run_cpu_chain(1, nb, 1 + 3)[1:]                       # B1: one channel on one core (BASELINE.json configs[0])
        per_pass = run_cpu_chain(cores, nb, 1 + reps)[1:]
        cpu_baseline = {"value": cores * nb * 32 * reps / sum(per_pass), "unit": "channel-samples/s", "cores": cores,
                        "kind": cpu_kind(), "sample": f"{cores} channels x {nb} blocks x {reps} passes, one process per channel",
                        "single_core": {"value": nb * 32 * 3 / sum(one), "unit": "channel-samples/s", "cores": 1,
                                        "sample": f"1 channel x {nb} blocks (10 s of signal) x 3 passes: BASELINE.json configs[0]"}}

    if rank == 0:
        peaks, which = measured_peaks()
        launch_ms = ms_per_step / max(1, launches / args.steps)
        tr = measured_traffic()
        launch_samples = nch * ns / max(1, launches / args.steps)
        roof = roofline_of(head, peaks, which, "ssb_narrow")
        roof.update({"traffic": tr["dram_bytes_per_channel_sample"] * launch_samples if tr else None,
                     "traffic_source": ("profiled constant, not measured in this run: " + tr["source"]) if tr else None,
                     "kernel": "rx_ssb_tc_kernel", "kernel_ms_per_launch": launch_ms, "algorithmic_bytes_per_launch": launch_samples * BYTES_PER_SAMPLE})
        roof["fp32"]["note"] = ("direct-form FLOP count of the chain (SURVEY.md 8d) against the FP32 FMA pipe, 148 SM x 128 lanes x 2 x 1.965 GHz; "
                                "282 of the 346 FLOP (the 83-tap decimator and the 199-tap Hilbert pair) run on the tensor cores as bf16-split Toeplitz GEMMs")
        line = {
            "metric": "channel-samples/s, full SSB RX chain", "value": value, "unit": "channel-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_LABEL, "data": "synthetic",
            "config": {"workload": f"{nch}-channel batched SSB (alternating USB path 35 / LSB path 38) RX chain per GPU, "
                                   f"{T} blocks (x32 samples) per channel per step (BASELINE.json configs[1])",
                       "channels_per_gpu": nch, "blocks_per_step": T, "parallelism": f"channels sharded over {world} GPU(s), no collective",
                       "l2": f"inputs larger than L2 ({2 * io_bytes / 2**20:.0f} MiB streamed per step)",
                       "warmup_note": f"{args.warmup} warm-up steps + {args.spinup_s} s of untimed launches of the same step before the timed region (sustained clocks)",
                       "generator": "uhsdr_b200/synth.py counter_block: integer counter-based, bit-identical on host (numpy) and device (torch)"},
            "roofline": roof,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "channel-samples/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes, "steps": e2e_steps,
                    "host_cpu_affinity_rank0": numa, "matches_device_path": e2e_match},
            "gpu_launches": launches, "clocks": clocks, "parity": head["parity"], "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--spinup-s", type=float, default=0.25, help="untimed spin-up under load after the warm-up steps, seconds")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=4096, help="channels per GPU (headline)")
    ap.add_argument("--blocks", type=int, default=1500, help="32-sample blocks per channel per step (headline)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--parity-channels", type=int, default=64, help="channels of every config compared with the oracle (0 = skip)")
    ap.add_argument("--parity-blocks", type=int, default=256)
    ap.add_argument("--other-blocks", type=int, default=400, help="blocks per step of the other BASELINE configs")
    ap.add_argument("--other-steps", type=int, default=3)
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
