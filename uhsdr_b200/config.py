"""ctypes mirror of uhsdr_chan_cfg_t / uhsdr_chan_status_t (include/uhsdr_b200.h).

The fields are the per-channel parameters the reference's block path reads from its globals
`ts`, `ads`, `agc_wdsp_conf`, `nr_params`, `sd`; defaults are the
reference's ui_configuration.c:60-220 values.
"""
from __future__ import annotations

import ctypes

DEMOD_USB, DEMOD_LSB, DEMOD_CW, DEMOD_AM, DEMOD_SAM, DEMOD_FM, DEMOD_DIGI = range(7)
SAM_SIDEBAND_BOTH, SAM_SIDEBAND_LSB, SAM_SIDEBAND_USB = range(3)
FREQ_IQ_CONV_OFF, FREQ_IQ_CONV_P6KHZ, FREQ_IQ_CONV_M6KHZ, FREQ_IQ_CONV_P12KHZ, FREQ_IQ_CONV_M12KHZ = range(5)
DSP_NR_ENABLE, DSP_NR_POSTAGC_ENABLE, DSP_NOTCH_ENABLE, DSP_NB_ENABLE, DSP_MNOTCH_ENABLE, DSP_MPEAK_ENABLE = (
    0x01, 0x02, 0x04, 0x08, 0x10, 0x20)
TX_FILTER_SOPRANO, TX_FILTER_TENOR, TX_FILTER_BASS = 1, 2, 3

TWINPEAKS_SAMPLING, TWINPEAKS_DONE, TWINPEAKS_WAIT, TWINPEAKS_UNCORRECTABLE, TWINPEAKS_CODEC_RESTART = range(5)

BLOCK_SIZE = 32
SAMPLE_RATE = 48000

_i32 = ctypes.c_int32
_f32 = ctypes.c_float


class ChanCfg(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("dmod_mode", _i32), ("filter_path", _i32), ("cw_lsb", _i32), ("digi_lsb", _i32),
        ("iq_freq_mode", _i32),
        ("iq_auto_correction", _i32), ("rx_adj_gain_i", _f32), ("rx_adj_gain_q", _f32),
        ("iq_phase_balance_rx", _f32),
        ("dsp_active", _i32), ("notch_frequency", _i32), ("peak_frequency", _i32),
        ("bass_gain", _i32), ("treble_gain", _i32), ("nr_strength", _i32), ("nb_setting", _i32),
        ("agc_mode", _i32), ("agc_slope", _i32), ("agc_hang_enable", _i32), ("agc_thresh", _i32),
        ("agc_hang_thresh", _i32), ("agc_hang_time", _i32), ("agc_tau_decay", _i32 * 6),
        ("agc_tau_hang_decay", _i32),
        ("sam_sideband", _i32), ("sam_fade_leveler", _i32), ("sam_pll_fmax", _i32),
        ("sam_zeta", _i32), ("sam_omegaN", _i32),
        ("fm_sql_threshold", _i32), ("fm_dev_5khz", _i32), ("fm_subaudible_tone_det_freq", _f32),
        ("nr_decimation_enable", _i32),
        ("spectrum_enable", _i32), ("spectrum_magnify", _i32), ("codec_gain_calc", _f32),
        ("tx_filter", _i32), ("tx_bass_gain", _i32), ("tx_treble_gain", _i32), ("tx_mic_gain", _i32),
        ("tx_comp_level", _i32), ("tx_alc_decay", _i32), ("tx_alc_postfilt_gain", _i32),
        ("tx_power_factor", _f32), ("tx_adj_gain_i", _f32), ("tx_adj_gain_q", _f32),
        ("iq_phase_balance_tx", _f32),
        ("notch_mu", _i32),
        ("fm_subaudible_tone_gen_freq", _f32), ("fm_tone_burst_mode", _i32),
    ]

    def copy(self) -> "ChanCfg":
        c = ChanCfg()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(self), ctypes.sizeof(ChanCfg))
        return c

    def replace(self, **kw) -> "ChanCfg":
        c = self.copy()
        for k, v in kw.items():
            if k == "agc_tau_decay":
                for i, x in enumerate(v):
                    c.agc_tau_decay[i] = x
            else:
                if not hasattr(c, k):
                    raise AttributeError(k)
                setattr(c, k, v)
        return c


class ChanStatus(ctypes.Structure):
    _fields_ = [
        ("adc_clip", _i32), ("adc_half_clip", _i32), ("adc_quarter_clip", _i32),
        ("agc_action", _i32), ("agc_hang_action", _i32),
        ("fm_squelched", _i32), ("fm_sql_avg", _f32),
        ("sam_carrier_freq_offset", _i32),
        ("iq_corr_c1", _f32), ("iq_corr_c2", _f32),
        ("tx_peak_audio", _f32), ("tx_alc_val", _f32),
        ("blocks_processed", ctypes.c_int64),
        ("twinpeaks_state", _i32), ("twinpeaks_restarts", _i32),
    ]


class SpectrumDisplayCfg(ctypes.Structure):
    """uhsdr_spectrum_display_cfg_t: ts.spectrum_db_scale / spectrum_agc_rate / spectrum_filter / dbm_constant and
    slayout.scope.w (ui_configuration.c:94,137,138,206)."""
    _fields_ = [("struct_size", _i32), ("spectrum_db_scale", _i32), ("spectrum_agc_rate", _i32), ("spectrum_filter", _i32),
                ("dbm_constant", _i32), ("scope_width", _i32)]


class SpectrumLevel(ctypes.Structure):
    _fields_ = [("dbm", _f32), ("dbmhz", _f32), ("display_offset", _f32)]


def default_spectrum_display_cfg(**kw) -> SpectrumDisplayCfg:
    """DB_DIV_10, SPECTRUM_SCOPE_AGC_DEFAULT 25, SPECTRUM_FILTER_DEFAULT 4, 480 display columns (the 480x320 layout)."""
    c = SpectrumDisplayCfg(ctypes.sizeof(SpectrumDisplayCfg), 3, 25, 4, 0, 480)
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def default_cfg(**kw) -> ChanCfg:
    """Reference defaults: USB, FilterPathInfo[35] (2.3 kHz LPF), -12 kHz translate, auto IQ
    correction, AGC mode 2 / slope 70 / thresh 20, bass +2 dB (ui_configuration.c:60-220)."""
    c = ChanCfg()
    c.struct_size = ctypes.sizeof(ChanCfg)
    c.dmod_mode = DEMOD_USB
    c.filter_path = 35
    c.iq_freq_mode = FREQ_IQ_CONV_M12KHZ
    c.iq_auto_correction = 1
    c.rx_adj_gain_i = 1.0
    c.rx_adj_gain_q = 1.0
    c.notch_frequency = 800
    c.peak_frequency = 750
    c.bass_gain = 2
    c.treble_gain = 0
    c.nr_strength = 160
    c.agc_mode = 2
    c.agc_slope = 70
    c.agc_hang_enable = 0
    c.agc_thresh = 20
    c.agc_hang_thresh = 45
    c.agc_hang_time = 500
    for i, v in enumerate((4000, 2000, 500, 250, 50, 1)):
        c.agc_tau_decay[i] = v
    c.agc_tau_hang_decay = 500
    c.sam_sideband = SAM_SIDEBAND_BOTH
    c.sam_fade_leveler = 1
    c.sam_pll_fmax = 2500
    c.sam_zeta = 65
    c.sam_omegaN = 250
    c.fm_sql_threshold = 12
    c.nr_decimation_enable = 1
    c.codec_gain_calc = 1.0
    c.tx_filter = TX_FILTER_SOPRANO
    c.tx_bass_gain = 4
    c.tx_treble_gain = 4
    c.tx_mic_gain = 15
    c.tx_comp_level = 2
    c.tx_alc_decay = 10
    c.tx_alc_postfilt_gain = 1
    c.tx_power_factor = 0.5
    c.tx_adj_gain_i = 1.0
    c.tx_adj_gain_q = 1.0
    c.notch_mu = 10
    return c.replace(**kw) if kw else c
