"""Pins the oracle.  The plain-C restatement (oracle/uhsdr_port.c) must reproduce, bit for bit, the
golden vectors that tests/golden/make_golden.py generated from the reference's own object code,
and -- where oracle/_ref is present -- the compiled reference itself on fresh seeded inputs."""
import os

import numpy as np
import pytest

from cases import SPECDISP_BLOCKS, SPECDISP_CASES, SPECDISP_REDRAWS, check_spectrum_display, FM_TONE_CASES, NB_CASES, NR_CASES, RX_CASES
from oracle import refchain
from oracle.port import PortChannel
from uhsdr_b200 import synth
from uhsdr_b200.config import default_cfg, default_spectrum_display_cfg

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "rx_golden.npz")


@pytest.fixture(scope="module")
def golden(built):
    return np.load(GOLDEN)


@pytest.mark.parametrize("label,kw,nblocks", RX_CASES, ids=[c[0] for c in RX_CASES])
def test_port_matches_golden_bit_exact(golden, label, kw, nblocks):
    cfg = default_cfg(**kw)
    iq = golden[f"{label}/iq"]
    with PortChannel(cfg) as p:
        audio, audio_f = p.rx(iq)
        st = p.status()
    assert np.array_equal(audio[:, 0], golden[f"{label}/audio_l"])
    assert np.array_equal(audio[:, 1], audio[:, 0])           # l == r in the OVI40 build (audio_driver.c:2868)
    assert np.array_equal(audio_f.view(np.uint32), golden[f"{label}/audio_f"].view(np.uint32))
    assert [st.adc_clip, st.agc_action, st.fm_squelched, st.sam_carrier_freq_offset] == list(golden[f"{label}/status"])
    # the run must not be trivially silent (FM opens its squelch at block 200)
    assert np.any(audio[:, 0] != 0)


def test_port_mute_and_reconfigure_sequence(golden):
    cfg_a, cfg_b = default_cfg(), default_cfg(filter_path=44, bass_gain=0)
    iq, mute = golden["seq_mute_reconf/iq"], golden["seq_mute_reconf/mute"]
    with PortChannel(cfg_a) as p:
        a1, f1 = p.rx(iq[: 80 * 32], mute[:80])
        p.reconfigure(cfg_b)
        a2, f2 = p.rx(iq[80 * 32:], mute[80:])
    got = np.concatenate([a1[:, 0], a2[:, 0]])
    assert np.array_equal(got, golden["seq_mute_reconf/audio_l"])
    assert np.all(got[40 * 32: 48 * 32] == 0)                 # external_mute -> zeros (audio_driver.c:2845-2853)


def test_port_block_granularity_invariance():
    """State carries across calls: one call of N blocks == N calls of one block."""
    cfg = default_cfg()
    iq = synth.rx_iq(cfg, 2, 64 * 32, seed=3)
    with PortChannel(cfg) as p:
        whole, _ = p.rx(iq)
    with PortChannel(cfg) as p:
        parts = [p.rx(iq[b * 32:(b + 1) * 32])[0] for b in range(64)]
    assert np.array_equal(whole, np.concatenate(parts))


def test_port_empty_call_and_zero_input():
    cfg = default_cfg()
    with PortChannel(cfg) as p:
        a, f = p.rx(np.zeros((0, 2), dtype=np.int32))
        assert a.shape == (0, 2)
        a, f = p.rx(np.zeros((32 * 8, 2), dtype=np.int32))
        assert np.all(a == 0) and np.all(f == 0)


def test_port_full_scale_input_sets_clip_flags():
    cfg = default_cfg()
    iq = np.full((32 * 4, 2), 2**31 - 1, dtype=np.int32)
    iq[::2] = -2**31 + 1
    with PortChannel(cfg) as p:
        p.rx(iq)
        st = p.status()
    assert st.adc_clip == 1 and st.adc_half_clip == 1 and st.adc_quarter_clip == 1


def test_port_rejects_unsupported():
    with pytest.raises(ValueError):
        PortChannel(default_cfg(spectrum_magnify=6))            # MAGNIFY_MAX is 5
    with pytest.raises(ValueError):
        PortChannel(default_cfg(filter_path=0))


@pytest.mark.skipif(not refchain.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("label,kw,nblocks", RX_CASES, ids=[c[0] for c in RX_CASES])
def test_port_matches_compiled_reference(built, label, kw, nblocks):
    cfg = default_cfg(**kw)
    iq = synth.rx_iq(cfg, 11, 2 * nblocks * 32, seed=999)
    with refchain.RefChannel(cfg) as r, PortChannel(cfg) as p:
        a_r, f_r = r.rx(iq)
        a_p, f_p = p.rx(iq)
    assert np.array_equal(a_r, a_p)
    assert np.array_equal(f_r.view(np.uint32), f_p.view(np.uint32))


# ---- spectral NR, spectrum FFT, TX ---------------------------------------------------------------
from cases import SPECTRUM_CASES, TX_CASES  # noqa: E402


@pytest.mark.parametrize("label,kw,nblocks", NR_CASES, ids=[c[0] for c in NR_CASES])
def test_port_nr_matches_golden_within_fft_rounding(golden, label, kw, nblocks):
    """The port's FFT is radix-2, the reference's radix-8: equal to float rounding, not bit-exact."""
    cfg = default_cfg(**kw)
    with PortChannel(cfg) as p:
        audio, audio_f = p.rx(golden[f"{label}/iq"])
    want = golden[f"{label}/audio_f"].astype(np.float64)
    # exact output latency: zeros until two processed frames are queued (audio_driver.c:2389-2417)
    thr = 1e-3 * np.max(np.abs(want))
    assert np.flatnonzero(np.abs(audio_f) > thr)[0] == np.flatnonzero(np.abs(want) > thr)[0]
    assert np.all(audio_f[: 32 * 32] == 0) and np.all(want[: 32 * 32] == 0)
    err = audio_f.astype(np.float64) - want
    assert np.max(np.abs(err)) <= 1e-5 * np.max(np.abs(want))
    assert 10 * np.log10(np.mean(want ** 2) / np.mean(err ** 2)) > 100.0


@pytest.mark.parametrize("label,kw,nblocks", NB_CASES, ids=[c[0] for c in NB_CASES])
def test_port_noise_blanker_matches_golden(golden, label, kw, nblocks):
    """LPC impulse blanker: alone it has no FFT on its path -> bit-exact; with the spectral NR behind it, FFT rounding.
    The golden inputs carry impulses, and the blanker must have repaired them (its output differs from the run without it)."""
    from uhsdr_b200.config import DSP_NB_ENABLE, DSP_NR_ENABLE
    cfg = default_cfg(**kw)
    iq = golden[f"{label}/iq"]
    with PortChannel(cfg) as p:
        audio, audio_f = p.rx(iq)
    want = golden[f"{label}/audio_f"]
    if cfg.dsp_active & DSP_NR_ENABLE:
        err = audio_f.astype(np.float64) - want.astype(np.float64)
        assert np.max(np.abs(err)) <= 1e-5 * np.max(np.abs(want))
    else:
        assert np.array_equal(audio[:, 0], golden[f"{label}/audio_l"])
        assert np.array_equal(audio_f.view(np.uint32), want.view(np.uint32))
    with PortChannel(default_cfg(**dict(kw, nb_setting=1))) as p:      # threshold 7.5 sigma x sqrt(LPC power): never fires, same latency
        _, delayed_f = p.rx(iq)
    assert np.count_nonzero(delayed_f != audio_f) > 500


@pytest.mark.parametrize("label,kw,nblocks,tone", FM_TONE_CASES, ids=[c[0] for c in FM_TONE_CASES])
def test_port_fm_subtone_detector_matches_golden(golden, label, kw, nblocks, tone):
    """The Goertzel tone detector gates the FM audio: bit-exact against the reference, open from block 800 on only when
    the tone is in the signal and the detector is tuned to it."""
    cfg = default_cfg(**kw)
    with PortChannel(cfg) as p:
        audio, audio_f = p.rx(golden[f"{label}/iq"])
    assert np.array_equal(audio[:, 0], golden[f"{label}/audio_l"])
    nz = np.flatnonzero(audio_f)
    if label == "fm_tone100_detected":
        assert nz[0] // 32 == 800
    else:
        assert len(nz) == 0


@pytest.mark.parametrize("label,kw", SPECTRUM_CASES, ids=[c[0] for c in SPECTRUM_CASES])
def test_port_spectrum_matches_golden(golden, label, kw):
    cfg = default_cfg(**kw)
    iq = golden[f"{label}/iq"]
    z = 1 << cfg.spectrum_magnify
    with PortChannel(cfg) as p:
        p.rx(iq[: 37 * z * 32])
        m1 = p.spectrum()
        p.rx(iq[37 * z * 32:])
        m2 = p.spectrum()
    for got, key in ((m1, "mags37"), (m2, "mags100")):
        want = golden[f"{label}/{key}"]
        assert np.max(np.abs(got - want)) <= 2e-6 * np.max(want)
        assert int(np.argmax(got)) == int(np.argmax(want))


@pytest.mark.parametrize("label,kw,dkw", SPECDISP_CASES, ids=[c[0] for c in SPECDISP_CASES])
def test_port_spectrum_display_matches_golden(label, kw, dkw):
    """UiSpectrum_RedrawSpectrum states 0-4 of the port against vectors from the reference's own ui_spectrum.c: bin averages
    to FFT rounding, dBm / dBm-per-Hz to 1e-3 dB, display columns and the sliding offset to 1e-4 of the display range."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spectrum_display_golden.npz"))
    cfg, dc = default_cfg(**kw), default_spectrum_display_cfg(**dkw)
    iq = g[f"{label}/iq"]
    n = SPECDISP_BLOCKS * (1 << cfg.spectrum_magnify) * 32
    with PortChannel(cfg) as p:
        for k in range(SPECDISP_REDRAWS):
            p.rx(iq[k * n:(k + 1) * n])
            mags, avg, disp, lvl = p.spectrum_display(dc)
            check_spectrum_display(g, label, k, mags, avg, disp, lvl, fft_tol=2e-6)


@pytest.mark.parametrize("label,kw,nblocks", TX_CASES, ids=[c[0] for c in TX_CASES])
def test_port_tx_matches_golden_bit_exact(golden, label, kw, nblocks):
    cfg = default_cfg(**kw)
    mic = np.zeros((nblocks * 32, 2), dtype=np.int32)
    mic[:, 0] = golden[f"{label}/mic"]
    with PortChannel(cfg) as p:
        iq, iq_f = p.tx(mic, golden[f"{label}/mute"])
        st = p.status()
    assert np.array_equal(iq, golden[f"{label}/iq"])
    assert np.array_equal(iq_f.view(np.uint32), golden[f"{label}/iq_f"].view(np.uint32))
    assert np.allclose([st.tx_alc_val, st.tx_peak_audio], golden[f"{label}/status"], rtol=0, atol=0)
    assert np.all(iq[(nblocks // 2) * 32:(nblocks // 2 + 5) * 32] == 0)     # muted blocks
    assert np.any(iq != 0)


def test_port_tx_rejects_modes_without_a_modulator():
    from uhsdr_b200.config import DEMOD_CW
    with PortChannel(default_cfg(dmod_mode=DEMOD_CW, filter_path=8)) as p:
        with pytest.raises(RuntimeError):
            p.tx(np.zeros((64, 2), dtype=np.int32))
