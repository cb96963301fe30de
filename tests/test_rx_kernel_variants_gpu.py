"""GPU parity tests for the kernel-selection logic of the shipping build: the tensor-core fused kernel
(rx_ssb_tc.cu, the default for narrow SSB/CW), the CUDA-core fused kernel (rx_ssb_fused.cu, used when
UHSDR_B200_NO_TC=1 or the output buffer is not 32-byte aligned) and the general kernel (call sizes that
are not a multiple of 4 blocks).  Every variant must meet north_star's tolerance against the oracle, and
a channel may move between them from call to call (the state records are shared)."""
import numpy as np
import pytest

from cases import RX_CASES
from conftest import oracle_channel
from test_rx_parity_gpu import check_tolerance, run_engine_float
from uhsdr_b200 import synth
from uhsdr_b200.config import DEMOD_LSB, default_cfg
from uhsdr_b200.engine import Engine

pytestmark = pytest.mark.gpu

FUSED_LABELS = ["usb_p35", "lsb_p38", "usb_p44_antialias", "cw_p8", "usb_agc_fast_hang", "usb_notch_peak_eq", "usb_manual_iq", "usb_agc_off"]
CASES = [c for c in RX_CASES if c[0] in FUSED_LABELS]


@pytest.mark.parametrize("label,kw,nblocks", CASES, ids=[c[0] for c in CASES])
def test_cuda_core_fused_kernel_within_tolerance(built, monkeypatch, label, kw, nblocks):
    monkeypatch.setenv("UHSDR_B200_NO_TC", "1")          # read at engine creation
    cfg = default_cfg(**kw)
    nb = 4 * nblocks
    iq = np.stack([synth.rx_iq(cfg, 200 + c, nb * 32, seed=44) for c in range(3)])
    with Engine(3) as eng:
        eng.configure(cfg)
        words, fl = run_engine_float(eng, iq)
    for c in range(3):
        with oracle_channel(cfg) as o:
            want_w, want_f = o.rx(iq[c])
        check_tolerance(fl[c], want_f, words[c, :, 0], want_w[:, 0], f"{label}/ch{c}")


def test_tensor_core_and_cuda_core_kernels_agree(built, monkeypatch):
    """Same inputs through both fused kernels: the float audio agrees to ~1e-5 of the peak."""
    import torch
    cfg_u, cfg_l = default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)
    nch, nb = 56, 128
    iq = np.stack([synth.rx_iq(cfg_u if c % 2 == 0 else cfg_l, c, nb * 32, seed=9) for c in range(nch)])
    outs = []
    for no_tc in ("0", "1"):
        monkeypatch.setenv("UHSDR_B200_NO_TC", no_tc)
        with Engine(nch) as eng:
            eng.configure(cfg_u, first=0, stride=2)
            eng.configure(cfg_l, first=1, stride=2)
            outs.append(run_engine_float(eng, iq)[1].astype(np.float64))
    err = np.abs(outs[0] - outs[1]).max() / np.abs(outs[1]).max()
    assert err < 5e-5, err
    assert np.isfinite(outs[0]).all()


def test_state_carries_between_kernels(built):
    """32 blocks (tensor-core kernel), 5 blocks (general kernel: not a multiple of 4), 27 blocks (general),
    64 blocks (tensor-core kernel again) on the same channels == one oracle run of 128 blocks."""
    cfg = default_cfg()
    nch, nb = 4, 128
    iq = np.stack([synth.rx_iq(cfg, 300 + c, nb * 32, seed=45) for c in range(nch)])
    parts_w, parts_f = [], []
    with Engine(nch) as eng:
        eng.configure(cfg)
        b0 = 0
        for n in (32, 5, 27, 64):
            w, f = run_engine_float(eng, iq[:, b0 * 32:(b0 + n) * 32])
            parts_w.append(w); parts_f.append(f)
            b0 += n
    words, fl = np.concatenate(parts_w, axis=1), np.concatenate(parts_f, axis=1)
    for c in range(nch):
        with oracle_channel(cfg) as o:
            want_w, want_f = o.rx(iq[c])
        check_tolerance(fl[c], want_f, words[c, :, 0], want_w[:, 0], f"mixed-kernels/ch{c}")


def test_unaligned_output_buffer_takes_the_cuda_core_kernel(built):
    """The tensor-core kernel stores 32-byte vectors; an output buffer that is only 16-byte aligned must still
    give correct results (engine falls back to the CUDA-core fused kernel)."""
    import torch
    cfg = default_cfg()
    nch, nb = 2, 64
    iq = np.stack([synth.rx_iq(cfg, 400 + c, nb * 32, seed=46) for c in range(nch)])
    dev = torch.device("cuda", 0)
    d_iq = torch.from_numpy(iq).to(dev)
    raw = torch.empty(nch * nb * 32 * 2 + 4, dtype=torch.int32, device=dev)
    d_out = raw[4:].view(nch, nb * 32, 2)                    # base + 16 bytes
    assert d_out.data_ptr() % 32 == 16
    with Engine(nch) as eng:
        eng.configure(cfg)
        eng.rx_device(d_iq, d_out, nb)
        eng.sync()
    got = d_out.cpu().numpy()
    for c in range(nch):
        with oracle_channel(cfg) as o:
            want_w, _ = o.rx(iq[c])
        diff = (got[c, :, 0].astype(np.int64) >> 16) - (want_w[:, 0].astype(np.int64) >> 16)
        assert np.max(np.abs(diff)) <= 1
        assert np.array_equal(got[c, :, 0], got[c, :, 1])


def test_tensor_core_path_with_spectrum_ring_and_nr_matches_the_split_path(built, monkeypatch):
    """Narrow SSB channels with the spectrum ring and / or the spectral NR run on the tensor-core kernel plus a spectrum tap kernel
    (+ NR kernel + serial phase 2).  Against the split general path (UHSDR_B200_NO_TCX=1) on the same inputs: audio within the float
    tolerance, identical NR latency, the spectrum of the last 512 samples equal to FFT rounding, same ring position; 30 channels
    (ragged CTA), a mute array, two calls (the second shorter than the 16-block ring)."""
    import torch
    from uhsdr_b200.config import DEMOD_LSB, DSP_NR_ENABLE
    cfgs = [default_cfg(spectrum_enable=1), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38, dsp_active=DSP_NR_ENABLE, spectrum_enable=1),
            default_cfg(dsp_active=DSP_NR_ENABLE)]
    nch, calls = 30, (96, 12)
    nb = sum(calls)
    iq = np.stack([synth.rx_iq(cfgs[c % 3], 500 + c, nb * 32, seed=51) for c in range(nch)])
    mute = np.zeros((nch, nb), dtype=np.uint8)
    mute[:, 40:44] = 1
    res = {}
    for no_tcx in ("1", "0"):
        monkeypatch.setenv("UHSDR_B200_NO_TCX", no_tcx)
        with Engine(nch) as eng:
            for c in range(nch):
                eng.configure(cfgs[c % 3], first=c, count=1)
            parts, pos = [], 0
            for n in calls:
                parts.append(run_engine_float(eng, iq[:, pos * 32:(pos + n) * 32], mute[:, pos:pos + n]))
                pos += n
            f = np.concatenate([p_[1] for p_ in parts], axis=1)
            mags = eng.spectrum()
            st = eng.status()
            res[no_tcx] = (f, mags, [s_.blocks_processed for s_ in st], eng.launch_count)
    f0, m0, b0, l0 = res["1"]
    f1, m1, b1, l1 = res["0"]
    assert b0 == b1 == [nb] * nch
    assert l1 != l0                                   # different kernels did run
    for c in range(nch):
        scale = np.max(np.abs(f0[c]))
        assert np.max(np.abs(f1[c] - f0[c])) <= 1e-4 * scale, c
        assert np.all(f1[c][40 * 32:44 * 32] == 0)
        if c % 3 != 2:                                 # channels with the spectrum ring
            assert np.max(np.abs(m1[c] - m0[c])) <= 1e-4 * np.max(m0[c]), c
            assert np.max(m0[c]) > 0
        if c % 3 != 0:                                 # NR: same latency
            thr = 1e-3 * scale
            assert np.flatnonzero(np.abs(f1[c]) > thr)[0] == np.flatnonzero(np.abs(f0[c]) > thr)[0], c


def test_unaligned_float_copy_buffer(built):
    """The optional float copy is stored as 16-byte vectors by the tensor-core kernel (8-byte by the CUDA-core fused kernel): a
    buffer that is only 4-byte aligned must take a kernel that stores it word by word, with the same result."""
    _unaligned_float_copy(default_cfg())
    _unaligned_float_copy(default_cfg(filter_path=48))      # split general path


def _unaligned_float_copy(cfg):
    import torch
    nch, nb = 2, 64
    iq = np.stack([synth.rx_iq(cfg, 410 + c, nb * 32, seed=47) for c in range(nch)])
    dev = torch.device("cuda", 0)
    d_iq = torch.from_numpy(iq).to(dev)
    res = []
    for off in (0, 1, 2):
        raw = torch.full((nch * nb * 32 + 4,), float("nan"), dtype=torch.float32, device=dev)
        d_f = raw[off:off + nch * nb * 32].view(nch, nb * 32)
        assert d_f.data_ptr() % 16 == 4 * off
        d_out = torch.empty_like(d_iq)
        with Engine(nch) as eng:
            eng.configure(cfg)
            eng.rx_device(d_iq, d_out, nb, audio_f_dev=d_f)
            eng.sync()
        assert bool(raw[:off].isnan().all()) and bool(raw[off + nch * nb * 32:].isnan().all())
        res.append((d_out.cpu().numpy(), d_f.cpu().numpy()))
    for w, f in res[1:]:
        assert np.max(np.abs((w[:, :, 0].astype(np.int64) >> 16) - (res[0][0][:, :, 0].astype(np.int64) >> 16))) <= 1
        assert np.max(np.abs(f - res[0][1])) <= 1e-4 * np.max(np.abs(res[0][1]))


def test_tensor_core_kernel_short_calls_carry_both_filter_histories(built):
    """The tensor-core kernel feeds the decimator / Hilbert histories through virtual steps and, for calls shorter than
    the 198-sample Hilbert history (fewer than 7 steps of 4 blocks), moves the part of the old history that survives.
    A run cut into calls of 1..6 steps must equal the one-call run bit for bit, and both must match the oracle."""
    cfg_u, cfg_l = default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)
    sizes = [4, 8, 4, 12, 16, 20, 24, 4, 28, 8]            # blocks per call (all multiples of 4: tensor-core kernel)
    nb, nch = sum(sizes), 30                                 # 30 channels: two CTAs, the second one ragged
    cfgs = [cfg_u if c % 3 else cfg_l for c in range(nch)]
    iq = np.stack([synth.rx_iq(cfgs[c], 500 + c, nb * 32, seed=46) for c in range(nch)])
    outs = []
    for cut in (False, True):
        with Engine(nch) as eng:
            for c in range(nch):
                eng.configure(cfgs[c], first=c, count=1)
            if not cut:
                outs.append(run_engine_float(eng, iq))
            else:
                ws, fs, pos = [], [], 0
                for n in sizes:
                    w, f = run_engine_float(eng, iq[:, pos * 32:(pos + n) * 32])
                    ws.append(w); fs.append(f); pos += n
                outs.append((np.concatenate(ws, axis=1), np.concatenate(fs, axis=1)))
    (w1, f1), (w2, f2) = outs
    assert np.array_equal(w1, w2)
    assert np.array_equal(f1.view(np.uint32), f2.view(np.uint32))
    for c in (0, 1, 27, 28, 29):
        with oracle_channel(cfgs[c]) as o:
            want_w, want_f = o.rx(iq[c])
        check_tolerance(f1[c], want_f, w1[c, :, 0], want_w[:, 0], f"short_calls/ch{c}")


def test_tensor_core_kernel_mute_and_float_copy(built):
    """external_mute through the tensor-core kernel: muted blocks are zeros, every filter state still advances
    (audio_driver.c:2845-2853), and the run without a mute array / float copy (the lean output path) gives the same words."""
    cfg = default_cfg()
    nch, nb = 5, 64
    iq = np.stack([synth.rx_iq(cfg, 600 + c, nb * 32, seed=47) for c in range(nch)])
    mute = np.zeros((nch, nb), dtype=np.uint8)
    mute[:, 8:13] = 1
    mute[2, 40:] = 1
    with Engine(nch) as eng:
        eng.configure(cfg)
        w_m, f_m = run_engine_float(eng, iq, mute)
    with Engine(nch) as eng:
        eng.configure(cfg)
        w_plain = eng.rx(iq)                                 # host-buffer entry point: no mute, no float copy
    mask = np.repeat(mute.astype(bool), 32, axis=1)
    assert np.all(w_m[mask] == 0) and np.all(f_m[mask] == 0.0)
    assert np.array_equal(w_m[~mask], w_plain[~mask])


def test_tensor_core_kernel_long_call(built):
    """One call of 3000 blocks (2 s of signal, 750 pipeline steps: every ring / accumulator / mbarrier phase wraps
    hundreds of times) against the oracle, and against the same signal in two calls, bit for bit."""
    cfg_u, cfg_l = default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)
    nch, nb = 6, 3000
    cfgs = [cfg_u if c % 2 == 0 else cfg_l for c in range(nch)]
    iq = np.stack([synth.rx_iq(cfgs[c], 700 + c, nb * 32, seed=48) for c in range(nch)])
    outs = []
    for cut in (None, 1996):
        with Engine(nch) as eng:
            for c in range(nch):
                eng.configure(cfgs[c], first=c, count=1)
            if cut is None:
                outs.append(run_engine_float(eng, iq))
            else:
                w1, f1 = run_engine_float(eng, iq[:, : cut * 32])
                w2, f2 = run_engine_float(eng, iq[:, cut * 32:])
                outs.append((np.concatenate([w1, w2], axis=1), np.concatenate([f1, f2], axis=1)))
    (w1, f1), (w2, f2) = outs
    assert np.array_equal(w1, w2)
    for c in range(nch):
        with oracle_channel(cfgs[c]) as o:
            want_w, want_f = o.rx(iq[c])
        check_tolerance(f1[c], want_f, w1[c, :, 0], want_w[:, 0], f"long/ch{c}")


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("path_kw", [dict(), dict(filter_path=48), dict(dmod_mode=3, filter_path=70)], ids=["narrow_fused", "wide_split", "am_split"])
def test_twinpeaks_detector_state(built, exact, path_kw):
    """AudioDriver_RxHandleTwinpeaks (audio_driver.c:2173-2248) as a host-visible flag: a healthy I/Q pair ends in DONE after
    1050 blocks, a pair without mirror rejection (Q = I: phase error 90 degrees) requests a codec restart; re-arming starts over;
    the fourth request in a row reads UNCORRECTABLE.  Same state sequence as the oracle, on the fused and on the split path."""
    from uhsdr_b200.config import TWINPEAKS_CODEC_RESTART, TWINPEAKS_DONE, TWINPEAKS_UNCORRECTABLE, TWINPEAKS_WAIT
    cfg = default_cfg(**path_kw)
    nb = 1100
    good = synth.counter_block(np, [synth.kind_of(cfg)], [3], 0, nb * 32)[0]
    twin = good.copy()
    twin[:, 1] = twin[:, 0]
    iq = np.stack([good, twin, good])
    with oracle_channel(cfg) as og, oracle_channel(cfg) as ot, Engine(3, exact=exact) as eng:
        eng.configure(cfg)
        seq, want = [], []
        for k in range(0, nb, 100):
            eng.rx(iq[:, k * 32:(k + 100) * 32])
            og.rx(good[k * 32:(k + 100) * 32]); ot.rx(twin[k * 32:(k + 100) * 32])
            st = eng.status()
            seq.append((st[0].twinpeaks_state, st[1].twinpeaks_state, st[2].twinpeaks_state))
            want.append((og.status().twinpeaks_state, ot.status().twinpeaks_state, og.status().twinpeaks_state))
        assert seq == want
        assert seq[0] == (TWINPEAKS_WAIT,) * 3 and seq[-1] == (TWINPEAKS_DONE, TWINPEAKS_CODEC_RESTART, TWINPEAKS_DONE)
        assert eng.status()[1].twinpeaks_restarts == 1
        for rep in range(3):                      # three more failed attempts: the fourth in a row is final
            eng.twinpeaks_rearm(first=1, count=1)
            ot.twinpeaks_rearm()
            assert eng.status()[1].twinpeaks_state == TWINPEAKS_WAIT
            for k in range(0, nb, 100):
                eng.rx(iq[:, k * 32:(k + 100) * 32])
                ot.rx(twin[k * 32:(k + 100) * 32])
            assert eng.status()[1].twinpeaks_state == ot.status().twinpeaks_state
        assert eng.status()[1].twinpeaks_state == TWINPEAKS_UNCORRECTABLE
        assert eng.status()[0].twinpeaks_state == TWINPEAKS_DONE


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_clip_flags_with_clipping_input(built, exact):
    """ads.adc_quarter_clip / adc_half_clip / adc_clip (audio_driver.c:2662-2675) against the oracle, fused and split path:
    one channel per threshold region (below a quarter, above a quarter, above half, full scale)."""
    for kw in (dict(), dict(filter_path=48)):
        cfg = default_cfg(**kw)
        base = synth.counter_block(np, [synth.kind_of(cfg)], [5], 0, 64 * 32)[0].astype(np.int64)
        chans = []
        for peak in (900, 1500, 3000, 6000):                     # thresholds are 1024 / 2048 / 4096 in 16-bit units
            x = (base.astype(np.float64) / np.max(np.abs(base[:, 0])) * (peak * 65536.0)).astype(np.int64)
            chans.append(np.clip(x, -2**31, 2**31 - 1).astype(np.int32))
        iq = np.stack(chans)
        with Engine(4, exact=exact) as eng:
            eng.configure(cfg)
            eng.rx(iq)
            got = [(s.adc_quarter_clip, s.adc_half_clip, s.adc_clip) for s in eng.status()]
        want = []
        for c in range(4):
            with oracle_channel(cfg) as o:
                o.rx(iq[c])
                s = o.status()
                want.append((s.adc_quarter_clip, s.adc_half_clip, s.adc_clip))
        assert got == want
        assert want == [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)]


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_host_path_time_slices_match_one_device_call(built, monkeypatch, exact):
    """uhsdr_rx_process cuts a large host call into time slices on three streams (H2D | kernels | D2H).  Forced here with
    UHSDR_B200_SLICES on a call whose block count is not a multiple of the slice size (the last slice takes the general kernel),
    with a mute array: the exact build is bit-equal to a single device-resident call (which, 282 blocks not being a multiple
    of 4, runs on the general kernel throughout); the shipping build, whose fused and general kernels differ by rounding, within
    one 16-bit LSB."""
    import torch
    monkeypatch.setenv("UHSDR_B200_SLICES", "5")
    nch, nb = 12, 70 * 4 + 2
    cfgs = [default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38), default_cfg(filter_path=48)]
    iq = np.concatenate([synth.counter_block(np, [synth.kind_of(cfgs[c % 3])], [c], 0, nb * 32) for c in range(nch)])
    big = np.tile(iq, (128, 1, 1))[: 128 * nch]                     # >= 32 MB so that the slicing engages
    n = big.shape[0]
    mute = np.zeros((n, nb), dtype=np.uint8)
    mute[:, 33:41] = 1
    mute[::3, 200:203] = 1
    outs = []
    for host in (True, False):
        with Engine(n, exact=exact) as eng:
            for k in range(3):
                eng.configure(cfgs[k], first=k, stride=3)
            if host:
                outs.append(eng.rx(big, mute))
            else:
                dev = torch.device("cuda", 0)
                d_iq, d_m = torch.from_numpy(big).to(dev), torch.from_numpy(mute).to(dev)
                d_out = torch.empty_like(d_iq)
                eng.rx_device(d_iq, d_out, nb, mute_dev=d_m)
                eng.sync()
                outs.append(d_out.cpu().numpy())
    if exact:
        assert np.array_equal(outs[0], outs[1])
    else:
        d = (outs[0].astype(np.int64) >> 16) - (outs[1].astype(np.int64) >> 16)
        assert int(np.max(np.abs(d))) <= 1
    assert np.all(outs[0][:, 33 * 32:41 * 32] == 0)


def test_c_host_program_multi_handle(built):
    """examples/host.c (plain C, uhsdr_multi_* over the device list) against the Python binding on the same input: the multi
    handle over one device -- or over two slots of the same device, which exercises the per-device ranges and host threads --
    returns what one engine returns."""
    import ctypes
    import os
    import subprocess
    import torch
    from uhsdr_b200.engine import load_library
    from uhsdr_b200.tables import DEFAULT_BLOB
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "host")
    outs = []
    devs = ["0", "1"] if torch.cuda.device_count() > 1 else ["0", "0"]
    for dl in (["0"], devs):
        r = subprocess.run([exe, DEFAULT_BLOB, "37", "24"] + dl, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.splitlines()[1])
        assert "abi=2" in r.stdout and "ch0_blocks=24" in r.stdout
    assert outs[0] == outs[1]
    # the same through ctypes: multi handle with ragged ranges (37 channels over 3 slots) == single engine
    lib = load_library()
    blob = open(DEFAULT_BLOB, "rb").read()
    buf = ctypes.create_string_buffer(blob, len(blob))
    nch, nb = 37, 24
    cfgs = [default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)]
    iq = np.concatenate([synth.counter_block(np, [synth.kind_of(cfgs[c % 2])], [c], 0, nb * 32) for c in range(nch)])
    with Engine(nch) as eng:
        eng.configure(cfgs[0], first=0, stride=2)
        eng.configure(cfgs[1], first=1, stride=2)
        want = eng.rx(iq)
    m = ctypes.c_void_p()
    dl = (ctypes.c_int * 3)(0, 0, 0)
    assert lib.uhsdr_multi_create(ctypes.byref(m), nch, dl, 3, buf, len(blob)) == 0
    assert lib.uhsdr_multi_configure_channels_strided(m, 0, 19, 2, ctypes.byref(cfgs[0]), 1) == 0
    assert lib.uhsdr_multi_configure_channels_strided(m, 1, 18, 2, ctypes.byref(cfgs[1]), 1) == 0
    got = np.empty_like(iq)
    assert lib.uhsdr_multi_rx_process(m, iq.ctypes.data, got.ctypes.data, nb, None) == 0
    lib.uhsdr_multi_destroy(m)
    assert np.array_equal(got, want)


def test_every_kernel_is_repeatable_and_stays_inside_its_buffers(built):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer.txt), so the two things it would have looked for
    are checked directly: (1) races -- every kernel family (scripts/sanitize.py: ragged CTAs, short and odd call sizes, fused,
    split, NR, TX, spectrum) run three times from fresh state must give bit-identical output; (2) out-of-bounds writes --
    device-pointer calls into buffers with canary rows in front, behind and BETWEEN the channel rows leave every canary
    untouched."""
    import io
    import os
    import sys
    import contextlib
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    import sanitize
    runs = []
    for _ in range(3):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            old = sys.argv
            sys.argv = ["sanitize.py"]
            try:
                sanitize.main()
            finally:
                sys.argv = old
        runs.append(buf.getvalue())
    assert runs[0] == runs[1] == runs[2]
    assert runs[0].count("crc=") >= 10
    # canaries: 30 channels (ragged CTA), rows padded by 64 samples each, calls of 8 and 20 blocks, RX and TX
    dev = torch.device("cuda", 0)
    for cfgs, nb in (([default_cfg(), default_cfg(dmod_mode=DEMOD_LSB, filter_path=38)] * 15, 8), ([default_cfg(filter_path=48)] * 13, 20)):
        n = len(cfgs)
        iq = np.concatenate([synth.counter_block(np, [synth.kind_of(c)], [i], 0, nb * 32) for i, c in enumerate(cfgs)])
        mic = np.concatenate([synth.counter_block(np, [synth.KIND_MIC], [i], 0, nb * 32) for i in range(n)])
        CAN = 0x5A5A5A5A
        with Engine(n) as eng:
            for i, c in enumerate(cfgs):
                eng.configure(c, first=i, count=1)
            d_iq, d_mic = torch.from_numpy(iq).to(dev), torch.from_numpy(mic).to(dev)
            pad = 4096
            out = torch.full((pad + n * nb * 32 + pad, 2), CAN, dtype=torch.int32, device=dev)
            out_f = torch.full((pad + n * nb * 32 + pad,), float("nan"), dtype=torch.float32, device=dev)
            eng.rx_device(d_iq, out[pad:], nb, audio_f_dev=out_f[pad:])
            eng.sync()
            assert bool((out[:pad] == CAN).all()) and bool((out[pad + n * nb * 32:] == CAN).all())
            assert bool(out_f[:pad].isnan().all()) and bool(out_f[pad + n * nb * 32:].isnan().all())
            assert not bool(out_f[pad:pad + n * nb * 32].isnan().any())
            txo = torch.full((pad + n * nb * 32 + pad, 2), CAN, dtype=torch.int32, device=dev)
            eng.tx_device(d_mic, txo[pad:], nb)
            eng.sync()
            assert bool((txo[:pad] == CAN).all()) and bool((txo[pad + n * nb * 32:] == CAN).all())
