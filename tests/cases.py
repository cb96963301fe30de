"""Shared parity cases: (label, cfg kwargs, nblocks).  Every distinct signal-flow topology, every
demodulator and the configuration switches the reference reads on the block path."""
from uhsdr_b200.config import (DEMOD_AM, DEMOD_CW, DEMOD_DIGI, DEMOD_FM, DEMOD_LSB, DEMOD_SAM, DSP_MNOTCH_ENABLE,
                               DSP_MPEAK_ENABLE, DSP_NB_ENABLE, DSP_NOTCH_ENABLE, DSP_NR_ENABLE, FREQ_IQ_CONV_M6KHZ, FREQ_IQ_CONV_OFF, FREQ_IQ_CONV_P12KHZ,
                               FREQ_IQ_CONV_P6KHZ, SAM_SIDEBAND_LSB, SAM_SIDEBAND_USB)

RX_CASES = [
    ("usb_p35", dict(), 96),
    ("lsb_p38", dict(dmod_mode=DEMOD_LSB, filter_path=38), 96),
    ("usb_p44_antialias", dict(filter_path=44), 96),
    ("usb_p48_hilbert_first", dict(filter_path=48), 96),
    ("usb_p55_dec2", dict(filter_path=55), 96),
    ("lsb_p65_dec2_antialias", dict(dmod_mode=DEMOD_LSB, filter_path=65), 96),
    ("cw_p8", dict(dmod_mode=DEMOD_CW, filter_path=8), 96),
    ("cw_lsb_p16", dict(dmod_mode=DEMOD_CW, filter_path=16, cw_lsb=1), 96),
    ("digi_lsb_p38", dict(dmod_mode=DEMOD_DIGI, filter_path=38, digi_lsb=1), 96),
    ("am_p70", dict(dmod_mode=DEMOD_AM, filter_path=70), 96),
    ("am_p84_nofade", dict(dmod_mode=DEMOD_AM, filter_path=84, sam_fade_leveler=0), 96),
    ("sam_p72_both", dict(dmod_mode=DEMOD_SAM, filter_path=72), 96),
    ("sam_p72_usb", dict(dmod_mode=DEMOD_SAM, filter_path=72, sam_sideband=SAM_SIDEBAND_USB), 96),
    ("sam_p84_lsb", dict(dmod_mode=DEMOD_SAM, filter_path=84, sam_sideband=SAM_SIDEBAND_LSB), 96),
    ("fm_p2", dict(dmod_mode=DEMOD_FM, filter_path=2), 480),
    ("fm_p1_sql0_5k", dict(dmod_mode=DEMOD_FM, filter_path=1, fm_sql_threshold=0, fm_dev_5khz=1), 320),
    ("usb_p6k", dict(iq_freq_mode=FREQ_IQ_CONV_P6KHZ), 96),
    ("usb_m6k", dict(iq_freq_mode=FREQ_IQ_CONV_M6KHZ), 96),
    ("usb_p12k", dict(iq_freq_mode=FREQ_IQ_CONV_P12KHZ), 96),
    ("usb_no_translate", dict(iq_freq_mode=FREQ_IQ_CONV_OFF), 96),
    ("usb_manual_iq", dict(iq_auto_correction=0, rx_adj_gain_i=1.01, rx_adj_gain_q=0.99, iq_phase_balance_rx=-0.01), 96),
    ("usb_manual_iq_pos", dict(iq_auto_correction=0, rx_adj_gain_i=0.98, rx_adj_gain_q=1.02, iq_phase_balance_rx=0.02), 96),
    ("usb_agc_off", dict(agc_mode=5), 96),
    ("usb_agc_fast_hang", dict(agc_mode=4, agc_hang_enable=1), 192),
    ("usb_agc_long_hang", dict(agc_mode=0, agc_hang_enable=1, agc_thresh=40, agc_slope=40), 192),
    ("usb_notch_peak_eq", dict(dsp_active=DSP_MNOTCH_ENABLE | DSP_MPEAK_ENABLE, treble_gain=3, bass_gain=-4), 96),
    # LMS automatic notch (DSP_NOTCH_ENABLE, audio_driver.c:1746-1763): narrow SSB, wide SSB with a fast rate, AM with the slowest
    ("usb_p35_autonotch", dict(dsp_active=DSP_NOTCH_ENABLE), 192),
    ("lsb_p48_autonotch_mu35", dict(dmod_mode=DEMOD_LSB, filter_path=48, dsp_active=DSP_NOTCH_ENABLE, notch_mu=35), 192),
    ("am_p70_autonotch_mu0", dict(dmod_mode=DEMOD_AM, filter_path=70, dsp_active=DSP_NOTCH_ENABLE, notch_mu=0), 192),
]

NR_CASES = [
    ("usb_p35_nr_dec", dict(dsp_active=DSP_NR_ENABLE), 320),
    ("usb_p44_nr_nodec", dict(filter_path=44, dsp_active=DSP_NR_ENABLE), 224),
    ("usb_p35_nr_decoff", dict(dsp_active=DSP_NR_ENABLE, nr_decimation_enable=0, nr_strength=100), 224),
]

# LPC impulse noise blanker (DSP_NB_ENABLE, alt_noise_blanking audio_nr.c:2210-2539); inputs carry seeded impulses
# (synth.add_impulses, NB_IMPULSES of them).  Alone it is integer/float-exact (no FFT); with the spectral NR behind it
# the FFT rounding applies.  The thresholds (16 - nb_setting)/2 sigma x sqrt(LPC power) are low enough that the repair path
# runs many times in every case (the tonal test signal makes the LPC power, hence the threshold, large).
NB_IMPULSES = 24
NB_CASES = [
    ("usb_p35_nb12", dict(dsp_active=DSP_NB_ENABLE, nb_setting=12), 320),
    ("usb_p44_nb14_nodec", dict(filter_path=44, dsp_active=DSP_NB_ENABLE, nb_setting=14), 224),
    ("usb_p35_nb15_nr", dict(dsp_active=DSP_NB_ENABLE | DSP_NR_ENABLE, nb_setting=15), 320),
]

# FM sub-audible tone detection (3 x Goertzel, audio_driver.c:1665-1734): (label, cfg, blocks, tone in the signal / Hz, 0 = none).
# The audio gate opens at the second evaluation window (block 800) when the tone is there and the detector is tuned to it.
FM_TONE_CASES = [
    ("fm_tone100_detected", dict(dmod_mode=DEMOD_FM, filter_path=2, fm_subaudible_tone_det_freq=100.0), 1000, 100.0),
    ("fm_tone100_detector_at_88", dict(dmod_mode=DEMOD_FM, filter_path=2, fm_subaudible_tone_det_freq=88.5), 1000, 100.0),
    ("fm_no_tone", dict(dmod_mode=DEMOD_FM, filter_path=2, fm_subaudible_tone_det_freq=100.0), 900, 0.0),
]

SPECTRUM_CASES = [
    ("spec_usb_p35", dict(spectrum_enable=1)),
    ("spec_fm_gain", dict(spectrum_enable=1, codec_gain_calc=2.5, dmod_mode=DEMOD_FM, filter_path=2)),
    # zoom FFT (sd.magnify 1..5, AudioDriver_SpectrumZoomProcessSamples audio_driver.c:1860-1909): biquad low-pass + decimation after the translation
    ("spec_zoom2_usb", dict(spectrum_enable=1, spectrum_magnify=1)),
    ("spec_zoom8_p6k", dict(spectrum_enable=1, spectrum_magnify=3, iq_freq_mode=FREQ_IQ_CONV_P6KHZ)),
    ("spec_zoom32_am", dict(spectrum_enable=1, spectrum_magnify=5, dmod_mode=DEMOD_AM, filter_path=70)),
]

# UiSpectrum_RedrawSpectrum states 0-4 (ui_spectrum.c:1362-1487): (label, channel kw, display-settings kw).  Four redraws per
# case, 24 blocks apart, so that the bin averages and the sliding display offset carry state.
SPECDISP_CASES = [
    ("disp_usb", dict(spectrum_enable=1), dict()),
    ("disp_lsb_w320_5db", dict(spectrum_enable=1, dmod_mode=DEMOD_LSB, filter_path=38), dict(scope_width=320, spectrum_db_scale=1, spectrum_filter=2)),
    ("disp_am_zoom4", dict(spectrum_enable=1, spectrum_magnify=2, dmod_mode=DEMOD_AM, filter_path=70), dict(spectrum_agc_rate=50, dbm_constant=-7)),
    ("disp_sam_usb_2s", dict(spectrum_enable=1, dmod_mode=DEMOD_SAM, filter_path=72, sam_sideband=2), dict(spectrum_db_scale=7, spectrum_filter=20)),
    ("disp_fm_p6k_w256", dict(spectrum_enable=1, dmod_mode=DEMOD_FM, filter_path=2, iq_freq_mode=FREQ_IQ_CONV_P6KHZ), dict(scope_width=256, spectrum_filter=1)),
    ("disp_cw_lsb_notrans", dict(spectrum_enable=1, dmod_mode=DEMOD_CW, filter_path=8, cw_lsb=1, iq_freq_mode=0), dict(scope_width=479, spectrum_agc_rate=1)),
]
SPECDISP_REDRAWS, SPECDISP_BLOCKS = 4, 24

TX_CASES = [
    ("tx_usb", dict(), 160),
    ("tx_lsb", dict(dmod_mode=DEMOD_LSB), 160),
    ("tx_usb_p6k_bass_comp8", dict(iq_freq_mode=FREQ_IQ_CONV_P6KHZ, tx_filter=3, tx_comp_level=8, iq_phase_balance_tx=0.01, tx_adj_gain_i=0.97), 160),
    ("tx_lsb_notrans_tenor_nocomp", dict(dmod_mode=DEMOD_LSB, iq_freq_mode=FREQ_IQ_CONV_OFF, tx_filter=2, tx_comp_level=-1, iq_phase_balance_tx=-0.02, tx_power_factor=0.05), 160),
    ("tx_usb_custom_comp", dict(tx_comp_level=13, tx_alc_decay=3, tx_alc_postfilt_gain=9, tx_mic_gain=40, iq_freq_mode=FREQ_IQ_CONV_P12KHZ), 160),
    # AM modulator (TxProcessor_AM, tx_processor.c:736-800): both sidebands + carrier, default -12 kHz translation and +6 kHz NCO
    ("tx_am", dict(dmod_mode=DEMOD_AM, filter_path=70), 160),
    ("tx_am_p6k_comp10", dict(dmod_mode=DEMOD_AM, filter_path=70, iq_freq_mode=FREQ_IQ_CONV_P6KHZ, tx_comp_level=10, tx_mic_gain=30), 160),
    # FM modulator (TxProcessor_FM, tx_processor.c:534-589): pre-emphasis + NCO on the sine table, 2.5 and 5 kHz deviation
    ("tx_fm", dict(dmod_mode=DEMOD_FM, filter_path=2), 160),
    ("tx_fm_5k_p6k", dict(dmod_mode=DEMOD_FM, filter_path=2, fm_dev_5khz=1, iq_freq_mode=FREQ_IQ_CONV_P6KHZ, tx_mic_gain=60), 160),
    ("tx_fm_m6k", dict(dmod_mode=DEMOD_FM, filter_path=2, iq_freq_mode=FREQ_IQ_CONV_M6KHZ), 160),
    # FM transmit tones (tx_processor.c:554-564): sub-audible tone on the modulation; tone burst (which silences the former)
    ("tx_fm_subtone_88", dict(dmod_mode=DEMOD_FM, filter_path=2, fm_subaudible_tone_gen_freq=88.5), 160),
    ("tx_fm_burst_1750_5k", dict(dmod_mode=DEMOD_FM, filter_path=2, fm_dev_5khz=1, fm_subaudible_tone_gen_freq=100.0, fm_tone_burst_mode=1), 160),
]

# float-math libm differences (sincosf / atan2f / expf) rule out bit-exactness for these
LIBM_CASES = {"sam_p72_both", "sam_p72_usb", "sam_p84_lsb", "fm_p2", "fm_p1_sql0_5k"}


def check_spectrum_display(g, label, k, mags, avg, disp, lvl, fft_tol):
    """Redraw k of a SPECDISP case against the reference vectors.  The display values are pixel rows (0 .. ~150): 1e-4 of that
    range is 1.5e-2; the sliding offset integrates the minimum over the log of the weakest bins, the most rounding-sensitive
    number of the display, and is held to 2e-3."""
    import numpy as np
    wm, wa, wd, wl = g[f"{label}/mags{k}"], g[f"{label}/avg{k}"], g[f"{label}/disp{k}"], g[f"{label}/lvl{k}"]
    if mags is not None:
        assert np.max(np.abs(mags - wm)) <= fft_tol * np.max(wm), (label, k)
    assert np.max(np.abs(avg - wa)) <= fft_tol * np.max(wa), (label, k)
    assert disp.shape == wd.shape
    assert np.max(np.abs(disp - wd)) <= 1e-4 * max(float(np.max(np.abs(wd))), 100.0), (label, k, float(np.max(np.abs(disp - wd))))
    assert abs(lvl[0] - wl[0]) <= 1e-3 and abs(lvl[1] - wl[1]) <= 1e-3, (label, k, lvl, wl)
    assert abs(lvl[2] - wl[2]) <= 2e-3, (label, k, lvl, wl)
