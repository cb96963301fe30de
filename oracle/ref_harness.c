/*
 * ref_harness.c -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Drives the reference's own RX/TX block processors, compiled unmodified from /root/reference,
 * for ONE channel per loaded copy of the shared object (the reference keeps all DSP state in
 * file-scope and function-local statics, e.g. audio_driver.c:1915-1916,1976,1531,
 * audio_agc.c:88,579, freq_shift.c:277-283).  A fresh channel == a freshly loaded copy.
 *
 * The reference translation unit is #included so that its static entry point
 * AudioDriver_RxProcessor (audio_driver.c:2603) and static coefficient arrays are reachable;
 * nothing of it is copied into this repository.
 */
#include "audio_driver.c"   /* from /root/reference via -I, see oracle/Makefile */

#include "uhsdr_b200.h"
#include "ui_configuration.h"
#include "tx_processor.h"
#include "arm_const_structs.h"

static uhsdr_chan_cfg_t g_cfg;
static int g_inited = 0;

static void ref_apply_cfg(const uhsdr_chan_cfg_t *c)
{
    ts.dmod_mode = c->dmod_mode;
    ts.cw_lsb = c->cw_lsb;
    ts.digi_lsb = c->digi_lsb;
    ts.iq_freq_mode = c->iq_freq_mode;
    ts.iq_auto_correction = c->iq_auto_correction;
    ts.rx_adj_gain_var.i = c->rx_adj_gain_i;
    ts.rx_adj_gain_var.q = c->rx_adj_gain_q;
    ads.iq_phase_balance_rx = c->iq_phase_balance_rx;
    ts.dsp.active = (uint8_t)c->dsp_active;
    ts.dsp.notch_frequency = c->notch_frequency;
    ts.dsp.peak_frequency = c->peak_frequency;
    ts.dsp.bass_gain = c->bass_gain;
    ts.dsp.treble_gain = c->treble_gain;
    ts.dsp.nr_strength = (uint8_t)c->nr_strength;
    ts.dsp.nb_setting = (uint8_t)c->nb_setting;
    ts.dsp.notch_numtaps = DSP_NOTCH_NUMTAPS_DEFAULT;
    ts.dsp.notch_mu = (uint8_t)c->notch_mu;
    ts.dsp.notch_delaybuf_len = DSP_NOTCH_DELAYBUF_DEFAULT;

    agc_wdsp_conf.mode = (uint8_t)c->agc_mode;
    agc_wdsp_conf.slope = (uint8_t)c->agc_slope;
    agc_wdsp_conf.hang_enable = (uint8_t)c->agc_hang_enable;
    agc_wdsp_conf.thresh = c->agc_thresh;
    agc_wdsp_conf.hang_thresh = c->agc_hang_thresh;
    agc_wdsp_conf.hang_time = c->agc_hang_time;
    for (int i = 0; i < 6; i++) agc_wdsp_conf.tau_decay[i] = c->agc_tau_decay[i];
    agc_wdsp_conf.tau_hang_decay = c->agc_tau_hang_decay;
    agc_wdsp_conf.switch_mode = 1;

    ads.sam_sideband = (sam_sideband_t)c->sam_sideband;
    ads.fade_leveler = (uint8_t)c->sam_fade_leveler;
    ads.pll_fmax_int = c->sam_pll_fmax;
    ads.zeta_int = c->sam_zeta;
    ads.omegaN_int = c->sam_omegaN;

    ts.fm_sql_threshold = (uint8_t)c->fm_sql_threshold;
    if (c->fm_dev_5khz) ts.flags2 |= FLAGS2_FM_MODE_DEVIATION_5KHZ; else ts.flags2 &= ~FLAGS2_FM_MODE_DEVIATION_5KHZ;

    nr_params.NR_decimation_enable = c->nr_decimation_enable;

    sd.magnify = (uint8_t)c->spectrum_magnify;
    sd.fft_iq_len = c->spectrum_enable ? 1024 : 0;   /* 480x320 layout: 512-point FFT, ui_spectrum.c:975-979 */
    ads.codec_gain_calc = c->codec_gain_calc;

    ts.tx_filter = (uint8_t)c->tx_filter;
    ts.dsp.tx_bass_gain = c->tx_bass_gain;
    ts.dsp.tx_treble_gain = c->tx_treble_gain;
    ts.tx_comp_level = (int16_t)c->tx_comp_level;
    ts.alc_decay = c->tx_alc_decay;
    ts.alc_tx_postfilt_gain = c->tx_alc_postfilt_gain;
    ts.tx_gain[TX_AUDIO_MIC] = (uint8_t)c->tx_mic_gain;
    ts.tx_mic_gain_mult = ts.tx_gain[TX_AUDIO_MIC];   /* ui_driver.c / radio_management: mic gain multiplier == setting */
    ts.tx_power_factor = c->tx_power_factor;
    ts.flags1 &= ~FLAGS1_SSB_TX_FILTER_DISABLE;
    if (g_inited) AudioManagement_CalcTxCompLevel();   /* the UI re-derives ALC decay / post-filter gain on change */
    for (int t = 0; t < IQ_TRANS_NUM; t++) {
        ts.tx_adj_gain_var[t].i = c->tx_adj_gain_i;
        ts.tx_adj_gain_var[t].q = c->tx_adj_gain_q;
        ads.iq_phase_balance_tx[t] = c->iq_phase_balance_tx;
    }
}

static void ref_select_path(int path)
{
    /* AudioDriver_SetProcessingChain re-reads the "last used in mode" memory (audio_driver.c:1105,
     * audio_filter.c:1032-1035); pin every mode's memory to the requested path. */
    uint16_t fm = AudioFilter_GetFilterModeFromDemodMode(ts.dmod_mode);
    ts.filter_path = (uint8_t)path;
    ts.filter_path_mem[fm][0] = (uint8_t)path;
}

/* FM transmit tones: the UI calls these when the settings change (audio_management.c:301-305, :339-349) and keys
 * tone_burst_active for the length of the burst; the oracle keeps it keyed while the mode is non-zero. */
static void ref_apply_fm_tones(const uhsdr_chan_cfg_t *cfg);

/* One-time init for this copy of the library == firmware boot: configuration load
 * (ui_configuration.c defaults, overridden by cfg), AudioDriver_Init (uhsdr_main.c:444), then
 * AudioDriver_SetProcessingChain (audio_driver.c:1093). */
int ref_init(const uhsdr_chan_cfg_t *cfg)
{
    if (cfg == NULL || cfg->struct_size != sizeof(uhsdr_chan_cfg_t)) return -1;
    if (g_inited) return -2;        /* statics cannot be re-zeroed: load a fresh copy instead */
    g_cfg = *cfg;
    memset((void *)&ts, 0, sizeof(ts));
    ts.samp_rate = 48000;
    ts.txrx_mode = TRX_MODE_RX;
    ts.tx_audio_source = TX_AUDIO_MIC;
    ts.rx_iq_source = RX_IQ_CODEC;
    ts.rx_gain[RX_AUDIO_SPKR].value = 10;
    ts.rx_gain[RX_AUDIO_DIG].value = 31;
    ts.beep_frequency = DEFAULT_BEEP_FREQUENCY;
    ts.twinpeaks_tested = TWINPEAKS_WAIT;                         /* uhsdr_main.c:339 */
    ref_apply_cfg(cfg);
    AudioDriver_Init();
    nr_params.NR_decimation_enable = cfg->nr_decimation_enable;   /* NR_Init (audio_nr.c:88) forces it to true */
    if (cfg->fm_subaudible_tone_det_freq > 0.0f) AudioManagement_CalcSubaudibleDetFreq(cfg->fm_subaudible_tone_det_freq);
    ref_apply_fm_tones(cfg);
    ref_select_path(cfg->filter_path);
    AudioDriver_SetProcessingChain(ts.dmod_mode, true);
    if (ts.filter_path != cfg->filter_path) return -3;
    g_inited = 1;
    return 0;
}

/* Reconfigure an initialised channel (reference semantics: AudioDriver_SetProcessingChain again). */
int ref_reconfigure(const uhsdr_chan_cfg_t *cfg)
{
    if (!g_inited || cfg == NULL || cfg->struct_size != sizeof(uhsdr_chan_cfg_t)) return -1;
    g_cfg = *cfg;
    ref_apply_cfg(cfg);
    if (cfg->fm_subaudible_tone_det_freq > 0.0f) AudioManagement_CalcSubaudibleDetFreq(cfg->fm_subaudible_tone_det_freq);
    ref_apply_fm_tones(cfg);
    ref_select_path(cfg->filter_path);
    AudioDriver_SetProcessingChain(ts.dmod_mode, false);
    return ts.filter_path == cfg->filter_path ? 0 : -3;
}

/* nblocks calls of AudioDriver_RxProcessor (audio_driver.c:2603).  iq/audio: nblocks*32 samples of
 * {int32 l, int32 r}.  audio_f (optional): adb.a_buffer[1] before the int conversion
 * (audio_driver.c:2911).  mute (optional): external_mute per block.  The deferred noise-reduction
 * task (PendSV, ui_driver.c:7157-7178) is run once after every block: the oracle's fixed schedule. */
int ref_rx(const int32_t *iq, int32_t *audio, float *audio_f, int nblocks, const uint8_t *mute)
{
    if (!g_inited) return -1;
    for (int b = 0; b < nblocks; b++) {
        IqSample_t src[IQ_BLOCK_SIZE];
        AudioSample_t dst[IQ_BLOCK_SIZE];
        memcpy(src, iq + (size_t)b * 2 * IQ_BLOCK_SIZE, sizeof(src));
        AudioDriver_RxProcessor(src, dst, IQ_BLOCK_SIZE, mute ? (mute[b] != 0) : false);
        memcpy(audio + (size_t)b * 2 * IQ_BLOCK_SIZE, dst, sizeof(dst));
        if (audio_f) memcpy(audio_f + (size_t)b * IQ_BLOCK_SIZE, adb.a_buffer[1], sizeof(float) * IQ_BLOCK_SIZE);
        AudioNr_HandleNoiseReduction();
    }
    return 0;
}

/* nblocks calls of TxProcessor_Run (tx_processor.c:891), SSB voice branch.  mic: nblocks*32 x
 * {int32 l, int32 r} (microphone in l); iq: same shape (l = I, r = Q); iq_f (optional): the float
 * I/Q after TxProcessor_IqFinalProcessing's scaling and phase mix, before the int conversion. */
int ref_tx(const int32_t *mic, int32_t *iq, float *iq_f, int nblocks, const uint8_t *mute)
{
    if (!g_inited) return -1;
    if (ts.txrx_mode != TRX_MODE_TX) {
        ts.txrx_mode = TRX_MODE_TX;
        TxProcessor_PrepareRun();           /* audio_driver.c:3012-3015 */
    }
    for (int b = 0; b < nblocks; b++) {
        AudioSample_t src[IQ_BLOCK_SIZE], side[IQ_BLOCK_SIZE];
        IqSample_t dst[IQ_BLOCK_SIZE];
        memcpy(src, mic + (size_t)b * 2 * IQ_BLOCK_SIZE, sizeof(src));
        TxProcessor_Run(src, dst, side, IQ_BLOCK_SIZE, mute ? (mute[b] != 0) : false);
        memcpy(iq + (size_t)b * 2 * IQ_BLOCK_SIZE, dst, sizeof(dst));
        if (iq_f) {
            for (int i = 0; i < IQ_BLOCK_SIZE; i++) {
                iq_f[((size_t)b * IQ_BLOCK_SIZE + i) * 2] = adb.iq_buf.i_buffer[i];
                iq_f[((size_t)b * IQ_BLOCK_SIZE + i) * 2 + 1] = adb.iq_buf.q_buffer[i];
            }
        }
    }
    return 0;
}

/* UiSpectrum_RedrawSpectrum states 0-2 (ui_spectrum.c:1362-1390) for the 480x320 layout
 * (fft_iq_len 1024, 512-point FFT, ui_spectrum.c:975-979), restated around the reference's own
 * CMSIS routines because ui_spectrum.c itself drags in the LCD stack.  `window` is von_Hann_1024
 * (ui_spectrum.c:362, a function-local constant; the caller passes it from the table blob). */
int ref_spectrum(const float *window, float *mags)
{
    static float samples[1024];
    if (!g_inited || sd.fft_iq_len != 1024) return -1;
    arm_copy_f32(&sd.FFT_RingBuffer[sd.samp_ptr], &samples[0], sd.fft_iq_len - sd.samp_ptr);
    arm_copy_f32(&sd.FFT_RingBuffer[0], &samples[sd.fft_iq_len - sd.samp_ptr], sd.samp_ptr);
    for (int i = 0; i < 1024; i++) samples[i] *= window[i];                 /* :409-413 */
    float32_t gcalc = 1.0 / ads.codec_gain_calc;                            /* :438-439 */
    arm_scale_f32(samples, gcalc, samples, 1024);
    arm_cfft_f32(&arm_cfft_sR_f32_len512, samples, 0, 1);                   /* :1383 */
    arm_cmplx_mag_f32(samples, mags, 512);                                  /* :1389 */
    return 0;
}

int ref_get_status(uhsdr_chan_status_t *st)
{
    memset(st, 0, sizeof(*st));
    st->adc_clip = ads.adc_clip; st->adc_half_clip = ads.adc_half_clip; st->adc_quarter_clip = ads.adc_quarter_clip;
    st->agc_action = agc_wdsp_conf.action; st->agc_hang_action = agc_wdsp_conf.hang_action;
    st->fm_squelched = ads.fm_conf.squelched; st->fm_sql_avg = ads.fm_conf.sql_avg;
    st->sam_carrier_freq_offset = ads.carrier_freq_offset;
    st->iq_corr_c1 = adb.iq_corr.M_c1; st->iq_corr_c2 = adb.iq_corr.M_c2;
    st->tx_peak_audio = ads.peak_audio; st->tx_alc_val = ads.alc_val;
    st->twinpeaks_state = ts.twinpeaks_tested;                    /* codec_restarts is a function-local static: not reachable */
    return 0;
}

/* Debug taps for the restatement's development: current I/Q scratch and audio scratch. */
const float *ref_tap_iq_i(void) { return adb.iq_buf.i_buffer; }
const float *ref_tap_iq_q(void) { return adb.iq_buf.q_buffer; }
const float *ref_tap_a0(void) { return adb.a_buffer[0]; }

static void ref_apply_fm_tones(const uhsdr_chan_cfg_t *cfg)
{
    AudioManagement_CalcSubaudibleGenFreq(cfg->fm_subaudible_tone_gen_freq > 0.0f ? cfg->fm_subaudible_tone_gen_freq : 0.0f);
    ts.fm_tone_burst_mode = (uint8_t)cfg->fm_tone_burst_mode;
    AudioManagement_LoadToneBurstMode();
    ads.fm_conf.tone_burst_active = cfg->fm_tone_burst_mode != 0;
}

/* what the firmware's main loop does after TWINPEAKS_CODEC_RESTART (ui_driver.c:7422-7425), minus the codec */
int ref_twinpeaks_rearm(void) { ts.twinpeaks_tested = TWINPEAKS_WAIT; return 0; }
