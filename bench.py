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
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except OSError:
        return None


def channel_cfg(ch: int):
    from uhsdr_b200.config import DEMOD_LSB, DEMOD_USB, default_cfg
    return default_cfg(dmod_mode=DEMOD_USB, filter_path=35) if ch % 2 == 0 else default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(idx, nblocks, reps, barrier, q):
    from oracle import refchain
    from oracle.port import PortChannel
    from uhsdr_b200 import synth
    cfg = channel_cfg(idx)
    iq = synth.rx_iq(cfg, idx, nblocks * 32)
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


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gen_iq_device(torch, nch, nsamples, device, ch0=0, seed=0x55485344):
    """Synthetic multi-tone + interferer + fading + AWGN I/Q generated on the device:
    int32 [nch, nsamples, 2], value = round(x * 2^16).  Even channels USB, odd channels LSB."""
    out = torch.empty((nch, nsamples, 2), dtype=torch.int32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(nsamples, dtype=torch.float64, device=device) / 48000.0
    two_pi = 2.0 * np.pi
    B = 64
    for c0 in range(0, nch, B):
        n = min(B, nch - c0)
        ch = torch.arange(ch0 + c0, ch0 + c0 + n, dtype=torch.float64, device=device)[:, None]
        sgn = torch.where((ch % 2) == 0, 1.0, -1.0)
        d = 3.0 * (ch % 64)
        fade = 10.0 ** ((6.0 * torch.sin(two_pi * 0.5 * t[None, :] + 0.1 * ch)) / 20.0)
        re = torch.zeros((n, nsamples), dtype=torch.float64, device=device)
        im = torch.zeros_like(re)
        for fa, amp in ((700.0, 3000.0), (1500.0, 2000.0), (2100.0, 1000.0)):
            ph = two_pi * ((12000.0 + sgn * (fa + d)) * t[None, :])
            re += amp * torch.cos(ph)
            im += amp * torch.sin(ph)
        re *= fade
        im *= fade
        ph = two_pi * ((12000.0 - sgn * 1500.0) * t[None, :])
        re += 3000.0 * torch.cos(ph)
        im += 3000.0 * torch.sin(ph)
        re += 100.0 * torch.randn((n, nsamples), dtype=torch.float64, device=device, generator=g)
        im += 100.0 * torch.randn((n, nsamples), dtype=torch.float64, device=device, generator=g)
        out[c0:c0 + n, :, 0] = torch.round(re * 65536.0).to(torch.int32)
        out[c0:c0 + n, :, 1] = torch.round(im * 65536.0).to(torch.int32)
        del re, im, fade, ph
    return out


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


def gpu_arm(args):
    import torch

    from uhsdr_b200.engine import Engine

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
    # host-side partitioning: rank r owns a contiguous range of the world*nch global channels
    from uhsdr_b200.partition import channel_range
    ch0, ch1 = channel_range(rank, world, world * nch)
    assert ch1 - ch0 == nch
    eng = Engine(nch, device=local_rank)
    from uhsdr_b200.config import DEMOD_LSB, default_cfg
    # even/odd channels alternate USB path 35 / LSB path 38
    cfg_usb, cfg_lsb = default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)
    if nch >= 2 and not args.uniform:
        first_usb = 0 if ch0 % 2 == 0 else 1
        eng.configure(cfg_usb, first=first_usb, stride=2)
        eng.configure(cfg_lsb, first=1 - first_usb, stride=2)
    else:
        eng.configure(cfg_usb)
    iq = gen_iq_device(torch, nch, ns, dev, ch0=ch0)
    audio = torch.empty_like(iq)
    torch.cuda.synchronize()

    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            eng.rx_device(iq, audio, T)
        barrier()
        launches0 = eng.launch_count
        t_begin = time.time()
        ev0.record(ext)
        for _ in range(args.steps):
            eng.rx_device(iq, audio, T)
        ev1.record(ext)
        barrier()
        launches = eng.launch_count - launches0
        # keep the same load running (untimed) until the clock sampler has seen MIN_LOAD_S of it
        while rank == 0 and time.time() - t_begin < ClockSampler.MIN_LOAD_S:
            for _ in range(4):
                eng.rx_device(iq, audio, T)
            torch.cuda.synchronize()
        clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
        barrier()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / args.steps
    total_samples = float(world) * nch * ns
    value = total_samples / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_iq = torch.empty((nch, ns, 2), dtype=torch.int32, pin_memory=True)
    h_audio = torch.empty((nch, ns, 2), dtype=torch.int32, pin_memory=True)
    h_iq.copy_(iq)
    torch.cuda.synchronize()
    lib, h = eng._lib, eng._h
    with torch.cuda.stream(ext):
        rc = lib.uhsdr_rx_process(h, h_iq.data_ptr(), h_audio.data_ptr(), T, None)   # warm-up (staging alloc)
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

    # ---- parity spot check against the oracle (not timed) -----------------------------------------
    parity = None
    if rank == 0 and args.parity_channels > 0:
        from oracle import refchain
        from oracle.port import PortChannel
        pc = min(nch, args.parity_channels)
        pb = min(T, 256) // 4 * 4
        sub = iq[:pc, : pb * 32].contiguous()
        sub_out = torch.empty_like(sub)
        with Engine(pc, device=local_rank) as peng:
            for c in range(pc):
                peng.configure(cfg_usb if (args.uniform or (ch0 + c) % 2 == 0) else cfg_lsb, first=c, count=1)
            peng.rx_device(sub, sub_out, pb)
            peng.sync()
        snrs = []
        for c in range(pc):
            cfg = cfg_usb if (args.uniform or (ch0 + c) % 2 == 0) else cfg_lsb
            x = sub[c].cpu().numpy()
            chan = refchain.RefChannel(cfg) if refchain.available() else PortChannel(cfg)
            want, _ = chan.rx(x)
            chan.close()
            got = sub_out[c, :, 0].cpu().numpy().astype(np.float64) / 65536.0
            ref = want[:, 0].astype(np.float64) / 65536.0
            snrs.append(10 * np.log10(np.mean(ref ** 2) / max(np.mean((got - ref) ** 2), 1e-30)))
        parity = {"channels": len(snrs), "blocks": pb, "min_snr_db_int16": float(min(snrs)), "oracle": cpu_kind()}

    # ---- cpu baseline (rank 0, N == 1 only) -------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        reps, nb = 10, 15000
        per_pass = run_cpu_chain(cores, nb, 1 + reps)[1:]
        cpu_baseline = {"value": cores * nb * 32 * reps / sum(per_pass), "unit": "channel-samples/s", "cores": cores,
                        "kind": cpu_kind(), "sample": f"{cores} channels x {nb} blocks x {reps} passes, one process per channel"}

    if rank == 0:
        peaks, which = measured_peaks()
        launch_ms = ms_per_step / max(1, launches / args.steps)
        per_gpu_samples = nch * ns
        achieved_gbs = per_gpu_samples * BYTES_PER_SAMPLE / (ms_per_step * 1e-3) / 1e9
        achieved_tf = per_gpu_samples * FLOP_PER_SAMPLE / (ms_per_step * 1e-3) / 1e12
        tr = measured_traffic()
        launch_samples = per_gpu_samples / max(1, launches / args.steps)
        traffic = tr["dram_bytes_per_channel_sample"] * launch_samples if tr else None
        line = {
            "metric": "channel-samples/s, full SSB RX chain", "value": value, "unit": "channel-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{nch}-channel batched SSB (alternating USB path 35 / LSB path 38) RX chain per GPU, "
                                   f"{T} blocks (x32 samples) per channel per step (BASELINE.json configs[1])",
                       "channels_per_gpu": nch, "blocks_per_step": T, "parallelism": f"channels sharded over {world} GPU(s), no collective",
                       "l2": f"inputs larger than L2 ({2 * io_bytes / 2**20:.0f} MiB streamed per step)"},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved_gbs / peaks["hbm_gbs"], "traffic": traffic, "peak_source": which,
                         "kernel": "rx_ssb_tc_kernel", "kernel_ms_per_launch": launch_ms,
                         "algorithmic_bytes_per_launch": launch_samples * BYTES_PER_SAMPLE,
                         "traffic_source": tr["source"] if tr else None,
                         "fp32": {"achieved_tflops": achieved_tf, "peak_tflops_nominal": FP32_PEAK_TFLOPS_NOMINAL,
                                  "frac": achieved_tf / FP32_PEAK_TFLOPS_NOMINAL, "flop_per_channel_sample": FLOP_PER_SAMPLE,
                                  "note": "direct-form FLOP count of the chain (SURVEY.md 8d) against the FP32 FMA pipe, 148 SM x 128 lanes x 2 x 1.965 GHz; "
                                          "282 of the 346 FLOP (the 83-tap decimator and the 199-tap Hilbert pair) run on the tensor cores as bf16-split Toeplitz GEMMs"}},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "channel-samples/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes, "steps": e2e_steps,
                    "host_cpu_affinity_rank0": numa},
            "gpu_launches": launches, "clocks": clocks, "parity": parity,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=4096, help="channels per GPU")
    ap.add_argument("--blocks", type=int, default=1500, help="32-sample blocks per channel per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--parity-channels", type=int, default=4)
    ap.add_argument("--uniform", action="store_true", help="all channels USB path 35")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
