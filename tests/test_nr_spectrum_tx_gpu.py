"""GPU parity: spectral noise reduction, spectrum-display FFT and the SSB transmit chain."""
import os

import numpy as np
import pytest

from cases import (NB_CASES, NB_IMPULSES, NR_CASES, SPECDISP_BLOCKS, SPECDISP_CASES, SPECDISP_REDRAWS, SPECTRUM_CASES, TX_CASES,
                   check_spectrum_display)
from conftest import oracle_channel
from test_rx_parity_gpu import check_tolerance, run_engine_float
from uhsdr_b200 import synth
from uhsdr_b200.config import DEMOD_AM, default_cfg, default_spectrum_display_cfg
from uhsdr_b200.engine import Engine, UhsdrError

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "rx_golden.npz")


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("label,kw,nblocks", NR_CASES, ids=[c[0] for c in NR_CASES])
def test_spectral_nr_within_tolerance(built, label, kw, nblocks, exact):
    cfg = default_cfg(**kw)
    nb = 3 * nblocks
    nch = 3
    iq = np.stack([synth.rx_iq(cfg, 20 + c, nb * 32, seed=17) for c in range(nch)])
    with Engine(nch, exact=exact) as eng:
        eng.configure(cfg)
        h = (nb // 3) * 32
        w1, f1 = run_engine_float(eng, iq[:, :h])
        w2, f2 = run_engine_float(eng, iq[:, h:])
    words, fl = np.concatenate([w1, w2], axis=1), np.concatenate([f1, f2], axis=1)
    for c in range(nch):
        with oracle_channel(cfg) as o:
            want_w, want_f = o.rx(iq[c])
        # the FIFO rule fixes the latency exactly (FFT round-off decides which of the first, nearly
        # silent samples are non-zero, so compare where the output first rises above 1e-3 of peak)
        thr = 1e-3 * np.max(np.abs(want_f))
        assert np.flatnonzero(np.abs(fl[c]) > thr)[0] == np.flatnonzero(np.abs(want_f) > thr)[0], label
        assert np.all(fl[c][: 32 * 32] == 0) and np.all(want_f[: 32 * 32] == 0)
        check_tolerance(fl[c], want_f, words[c, :, 0], want_w[:, 0], f"{label}/ch{c}")


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("label,kw,nblocks", NR_CASES[:2], ids=[c[0] for c in NR_CASES[:2]])
def test_nr_call_size_invariance(built, label, kw, nblocks, exact):
    """The NR kernel walks the frame interface in runs of blocks between FIFO events (input frame full, output frame used up, a
    frame processed).  Where a call ends must not matter: one call of N blocks and the same signal cut into calls of irregular sizes
    (1, 3, 7, 33, 64 ... blocks: runs cut short at every position relative to the events, kernels of different generations for the
    small calls) give the same bits."""
    cfg = default_cfg(**kw)
    nb = 480
    nch = 2
    iq = np.stack([synth.rx_iq(cfg, 60 + c, nb * 32, seed=23) for c in range(nch)])
    with Engine(nch, exact=exact) as eng:
        eng.configure(cfg)
        w_one, f_one = run_engine_float(eng, iq)
    sizes, left = [], nb
    for sz in [1, 3, 7, 33, 64, 5, 96, 2, 31, 128, 17]:
        if left <= 0:
            break
        sizes.append(min(sz, left)); left -= sizes[-1]
    if left > 0:
        sizes.append(left)
    with Engine(nch, exact=exact) as eng:
        eng.configure(cfg)
        parts, pos = [], 0
        for sz in sizes:
            parts.append(run_engine_float(eng, iq[:, pos * 32:(pos + sz) * 32]))
            pos += sz
    w_cut = np.concatenate([p[0] for p in parts], axis=1)
    f_cut = np.concatenate([p[1] for p in parts], axis=1)
    if exact:
        assert np.array_equal(f_cut.view(np.uint32), f_one.view(np.uint32)), label
        assert np.array_equal(w_cut, w_one), label
    else:
        # small calls take the general kernels (other FIR summation order): float tolerance, same latency
        assert np.max(np.abs(f_cut - f_one)) <= 1e-4 * np.max(np.abs(f_one)), label
        thr = 1e-3 * np.max(np.abs(f_one))
        for c in range(nch):
            assert np.flatnonzero(np.abs(f_cut[c]) > thr)[0] == np.flatnonzero(np.abs(f_one[c]) > thr)[0], label


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("label,kw,nblocks", NB_CASES, ids=[c[0] for c in NB_CASES])
def test_lpc_noise_blanker(built, label, kw, nblocks, exact):
    """alt_noise_blanking (audio_nr.c:2210-2539) on inputs with impulses, settings at which the repair path fires on most
    frames.  The blanker is a threshold decision followed by a replacement, so it keeps the reference's operation order in
    both builds: alone (no FFT on the path) the exact build must be bit-exact, and the fast build / the case with the
    spectral NR behind it must stay within the float tolerance -- one different decision would break it by far."""
    from uhsdr_b200.config import DSP_NR_ENABLE
    cfg = default_cfg(**kw)
    nb = 3 * nblocks
    nch = 3
    iq = np.stack([synth.add_impulses(synth.rx_iq(cfg, 40 + c, nb * 32, seed=19), 19 + c, 3 * NB_IMPULSES) for c in range(nch)])
    with Engine(nch, exact=exact) as eng:
        eng.configure(cfg)
        h = (nb // 3) * 32
        w1, f1 = run_engine_float(eng, iq[:, :h])
        w2, f2 = run_engine_float(eng, iq[:, h:])
    words, fl = np.concatenate([w1, w2], axis=1), np.concatenate([f1, f2], axis=1)
    for c in range(nch):
        with oracle_channel(cfg) as o:
            want_w, want_f = o.rx(iq[c])
        if exact and not (cfg.dsp_active & DSP_NR_ENABLE):
            assert np.array_equal(words[c], want_w), (label, c)
            assert np.array_equal(fl[c].view(np.uint32), want_f.view(np.uint32)), (label, c)
        else:
            check_tolerance(fl[c], want_f, words[c, :, 0], want_w[:, 0], f"{label}/ch{c}")


@pytest.mark.parametrize("label,kw", SPECTRUM_CASES, ids=[c[0] for c in SPECTRUM_CASES])
def test_spectrum_fft(built, label, kw):
    g = np.load(GOLDEN)
    cfg = default_cfg(**kw)
    iq = g[f"{label}/iq"]
    nch = 4
    batch = np.stack([iq] * nch)
    with Engine(nch) as eng:
        eng.configure(cfg)
        z = 1 << cfg.spectrum_magnify
        eng.rx(batch[:, : 37 * z * 32])
        m1 = eng.spectrum()
        eng.rx(batch[:, 37 * z * 32:])
        m2 = eng.spectrum(first=1, count=2)
    for got, key in ((m1[0], "mags37"), (m1[3], "mags37"), (m2[0], "mags100"), (m2[1], "mags100")):
        want = g[f"{label}/{key}"]
        assert np.max(np.abs(got - want)) <= 1e-4 * np.max(want), label
        assert int(np.argmax(got)) == int(np.argmax(want))


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("label,kw,dkw", SPECDISP_CASES, ids=[c[0] for c in SPECDISP_CASES])
def test_spectrum_display_states_3_4(built, label, kw, dkw, exact):
    """Bin averaging, dBm / dBm-per-Hz, log scaling, width rescaling and the sliding display offset (ui_spectrum.c:1432-1487,
    :1990-2122) against vectors from the reference's own ui_spectrum.c; two channels of a batch, a sub-range call in between."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spectrum_display_golden.npz"))
    cfg, dc = default_cfg(**kw), default_spectrum_display_cfg(**dkw)
    iq = g[f"{label}/iq"]
    n = SPECDISP_BLOCKS * (1 << cfg.spectrum_magnify) * 32
    nch = 3
    batch = np.stack([iq] * nch)
    with Engine(nch, exact=exact) as eng:
        eng.configure(cfg)
        for k in range(SPECDISP_REDRAWS):
            eng.rx(batch[:, k * n:(k + 1) * n])
            if k == 2:          # channel 1 alone first, then the other two: per-channel state, any sub-range
                d1, l1, a1 = eng.spectrum_display(dc, first=1, count=1)
                d0, l0, a0 = eng.spectrum_display(dc, first=0, count=1)
                d2, l2, a2 = eng.spectrum_display(dc, first=2, count=1)
                disp, lvl, avg = np.concatenate([d0, d1, d2]), np.concatenate([l0, l1, l2]), np.concatenate([a0, a1, a2])
            else:
                disp, lvl, avg = eng.spectrum_display(dc)
            for c in range(nch):
                check_spectrum_display(g, label, k, None, avg[c], disp[c], lvl[c], fft_tol=1e-5)


def test_spectrum_requires_enable(built):
    with Engine(2) as eng:
        eng.configure(default_cfg())
        with pytest.raises(UhsdrError):
            eng.spectrum()


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("label,kw,nblocks", TX_CASES, ids=[c[0] for c in TX_CASES])
def test_tx_ssb(built, label, kw, nblocks, exact):
    import torch
    g = np.load(GOLDEN)
    cfg = default_cfg(**kw)
    mic = np.zeros((nblocks * 32, 2), dtype=np.int32)
    mic[:, 0] = g[f"{label}/mic"]
    mute = g[f"{label}/mute"]
    nch = 3
    with Engine(nch, exact=exact) as eng:
        eng.configure(cfg)
        dev = torch.device("cuda", 0)
        d_mic = torch.from_numpy(np.stack([mic] * nch)).to(dev)
        d_mute = torch.from_numpy(np.stack([mute] * nch)).to(dev)
        d_iq = torch.empty_like(d_mic)
        d_f = torch.empty(d_mic.shape, dtype=torch.float32, device=dev)
        h = (nblocks // 2 + 3)
        a, b = d_mic[:, : h * 32].contiguous(), d_mic[:, h * 32:].contiguous()
        oa, ob = torch.empty_like(a), torch.empty_like(b)
        fa, fb = torch.empty(a.shape, dtype=torch.float32, device=dev), torch.empty(b.shape, dtype=torch.float32, device=dev)
        ma, mb = d_mute[:, :h].contiguous(), d_mute[:, h:].contiguous()   # keep alive: the calls are asynchronous
        eng.tx_device(a, oa, h, iq_f_dev=fa, mute_dev=ma)
        eng.tx_device(b, ob, nblocks - h, iq_f_dev=fb, mute_dev=mb)
        eng.sync()
        iq = torch.cat([oa, ob], dim=1).cpu().numpy()
        iq_f = torch.cat([fa, fb], dim=1).cpu().numpy()
        st = eng.status()
    want, want_f = g[f"{label}/iq"], g[f"{label}/iq_f"]
    for c in range(nch):
        if exact:
            assert np.array_equal(iq[c], want), (label, c)
            assert np.array_equal(iq_f[c].view(np.uint32), want_f.view(np.uint32))
        else:
            err = iq_f[c].astype(np.float64) - want_f
            assert np.max(np.abs(err)) <= 1e-4 * np.max(np.abs(want_f)), label
            assert 10 * np.log10(np.mean(want_f.astype(np.float64) ** 2) / max(np.mean(err ** 2), 1e-30)) >= 90.0
            assert np.max(np.abs(iq[c].astype(np.int64) - want.astype(np.int64))) <= max(1.0, 1e-4 * np.max(np.abs(want_f)))
        assert abs(st[c].tx_alc_val - g[f"{label}/status"][0]) <= 1e-5
    # host-buffer entry point
    with Engine(1, exact=True) as eng:
        eng.configure(cfg)
        got = eng.tx(mic[None], mute[None])
    assert np.array_equal(got[0], want)


def test_tx_modes_without_a_modulator_are_rejected(built):
    from uhsdr_b200.config import DEMOD_CW
    with Engine(1) as eng:
        eng.configure(default_cfg(dmod_mode=DEMOD_CW, filter_path=8))
        with pytest.raises(UhsdrError) as ei:
            eng.tx(np.zeros((1, 64, 2), dtype=np.int32))
        assert ei.value.code == -5


def test_rx_after_tx_shares_the_translate_oscillator(built):
    """FreqShift keeps one NCO for RX and TX (freq_shift.c:277-283 statics): RX after TX at +6 kHz
    continues the oscillator where TX left it."""
    cfg = default_cfg(iq_freq_mode=1)       # FREQ_IQ_CONV_P6KHZ
    mic = synth.tx_mic(1, 64 * 32)
    iq = synth.rx_iq(cfg, 1, 96 * 32, seed=4)
    with Engine(1, exact=True) as eng:
        eng.configure(cfg)
        eng.tx(mic[None])
        got = eng.rx(iq[None])
    with oracle_channel(cfg) as o:
        o.tx(mic)
        want, _ = o.rx(iq)
    assert np.array_equal(got[0], want)
