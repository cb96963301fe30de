// uhsdr_dev.h -- per-channel parameter and state records shared by the host setup code and the
// CUDA kernels.  One record per channel; the reference keeps the same quantities in file-scope
// statics (audio_driver.c:429-430 `ads`/`adb`, audio_agc.c:88 `agc_wdsp`, audio_driver.c:1531
// `fm_data`, :1976 `sam_data`, freq_shift.c:277-283, CMSIS instance pState arrays).
#pragma once
#include <stdint.h>

namespace uhsdr {

constexpr int BLK = 32;            // IQ_BLOCK_SIZE, uhsdr_board_config.h:217
constexpr int CHUNK_BLOCKS = 4;    // blocks handled per inner iteration of the generic kernel
constexpr int CHUNK = BLK * CHUNK_BLOCKS;
constexpr int H1 = 96;             // stage-1 history slots @48k (>= 88)
constexpr int H2 = 200;            // stage-2 history slots (>= 198)
constexpr int AGC_RB = 192;        // AGC_WDSP_RB_SIZE, audio_agc.c:19
constexpr int MAX_LAT = 12;
constexpr int INTERP_HIST = 8;     // phaseLength-1 <= 7, arm_fir_interpolate_f32.c

// signal-flow topologies of AudioDriver_RxProcessor (audio_driver.c:2718-2829)
enum Topo : int {
    TOPO_SSB_DEC_FIRST = 1,   // use_decimatedIQ: decimate I,Q -> Hilbert pair at the decimated rate -> I+-Q
    TOPO_SSB_HIL_FIRST = 2,   // Hilbert pair @48k -> I+-Q -> decimate the audio
    TOPO_AM_SAM = 3,          // decimate I,Q with the AM low-pass -> AudioDriver_DemodSAM
    TOPO_FM = 4               // FIR pair @48k -> AudioDriver_DemodFM, no decimation
};

struct LatticeP { int n; int k_off; int v_off; };   // offsets (in floats) into the coefficient pool

// AudioAgc_SetupAgcWdsp results, audio_agc.c:126-339
struct AgcP {
    int mode, hang_enable, attack_buffsize, remove_dc;
    float sample_rate, hangtime;
    float fixed_gain, attack_mult, decay_mult, fast_decay_mult, fast_backmult, onemfast_backmult;
    float out_target, min_volts, slope_constant, inv_max_input, pop_ratio;
    float hang_level, hang_backmult, onemhang_backmult, hang_decay_mult;
};

struct ChanParams {
    int configured;
    int mode;              // DemodModes_t
    int topo;
    int lsb;               // RadioManagement_LSBActive for SSB/CW/DIGI
    int M;                 // decimation factor (ads.decimation_rate)
    int decimated_freq;
    // IQ correction, audio_driver.c:2254-2316
    int iq_auto;
    float adj_i, adj_q, phase_bal;
    // FreqShift, freq_shift.c:275-331
    int shift_kind;        // 0 none, 1 quarter-fs sign/swap, 2 recursive NCO
    int shift_down;        // dir = shift > 0
    int shift_freq;        // |shift| in Hz (conversion_freq)
    float osc_cos, osc_sin;
    // stage 1 (48 ksps input): decimator (DEC_FIRST, AM_SAM) or Hilbert/LPF pair (HIL_FIRST, FM)
    int s1_ntaps, s1_ci, s1_cq, s1_M;
    // stage 2: Hilbert pair at the decimated rate (DEC_FIRST) or audio decimator (HIL_FIRST)
    int s2_ntaps, s2_ci, s2_cq, s2_M;
    LatticeP pre, aa, sql;
    int interp_L, interp_plen, interp_c;
    float bq1[4][5];
    float bq2[5];
    float scale_gain;      // audio_driver.c:2513-2521
    AgcP agc;
    // SAM, audio_driver.c:709-745
    int sam_sideband, fade_leveler;
    float sam_omega_min, sam_omega_max, sam_g1, sam_g2, sam_mtauR, sam_onem_mtauR, sam_mtauI, sam_onem_mtauI;
    int sam_c0, sam_c1;    // pool offsets of demod_sam_const.c0/.c1
    // FM
    int fm_sql_threshold;
    int fm_tone_det;        // ads.fm_conf.subaudible_tone_det_freq != 0
    float fm_gz_r[3], fm_gz_cos[3], fm_gz_sin[3];   // AudioFilter_CalcGoertzel for 1.04 f, 0.95 f, f over 400 blocks (audio_management.c:311-326)
    float fm_scaling;
    int fm_translate_on;
    // spectral NR
    int nr_enable;         // the NR frame interface runs: sampleRateDecim == 12000 && (is_dsp_nr() || is_dsp_nb_active()), audio_driver.c:2501
    int nr_spectral;       // is_dsp_nr(): spectral_noise_reduction_3 on every frame
    int nb_enable;         // is_dsp_nb_active(): alt_noise_blanking on every frame (before the spectral NR), audio_nr.c:362-365
    int nb_level;          // 16 - ts.dsp.nb_setting (threshold in half standard deviations, audio_nr.c:2414)
    int nr_decim;          // nr_params.NR_decimation_active
    float nr_alpha;
    int nr_vad_low, nr_vad_high;
    int nr_dec_c, nr_int_c, nr_win_c;   // pool offsets
    int tw256_off, tw512_off;           // FFT twiddle tables in the pool
    float nr_xih1;                      // pow10f(NR2.asnr / 10), audio_nr.c:1886
    // spectrum
    int spectrum_enable;
    int zoom_m;            // sd.magnify 1..5 (0 = no zoom): AudioDriver_SpectrumZoomProcessSamples, audio_driver.c:1860-1909
    int zoom_bq_off;       // pool offset of mag_coeffs[zoom_m] (4 stages x {b0, b1, b2, a1, a2})
    int zoom_dec_off;      // pool offset of FirZoomFFTDecimate[zoom_m].pCoeffs (4 taps)
    float codec_gain_calc;
    // UiSpectrum_CalculateDBm (ui_spectrum.c:1990-2122): passband bins [dbm_lbin, dbm_ubin] of the frequency-ordered spectrum
    // and the bandwidth ((int)Ubin - (int)Lbin) * bin_BW the dBm/Hz figure is referred to
    int dbm_lbin, dbm_ubin;
    float dbm_span_hz;
    // LMS automatic notch (audio_driver.c:1165-1187, :2443-2456)
    int notch_enable;
    float notch_mu;        // log10f((ts.dsp.notch_mu + 1.0) / 1500.0 + 1.0)
};

constexpr int NOTCH_TAPS = 64;     // DSP_NOTCH_NUMTAPS_MIN == MAX, audio_driver.h:486-488
constexpr int NOTCH_DELAY = 128;   // DSP_NOTCH_BUFLEN_MIN == MAX, audio_driver.h:490-492

struct BiquadS { float x1, x2, y1, y2; };

struct ChanState {
    // iq_correction_data_t, audio_driver.h:125-135 (teta*_old, M_c1, M_c2)
    float teta1_old, teta2_old, teta3_old, M_c1, M_c2;
    // FreqShift_Approx oscillator vector, freq_shift.c:20-47
    float osc_vect_q, osc_vect_i;
    int conversion_freq;
    // FIR histories (oldest first, right-aligned in the H1/H2 slots)
    float s1_hist_i[H1], s1_hist_q[H1];
    float s2_hist_i[H2], s2_hist_q[H2];
    float interp_hist[INTERP_HIST];
    float pre_s[MAX_LAT], aa_s[MAX_LAT], sql_s[MAX_LAT];
    BiquadS bq1[4], bq2;
    // AGC, audio_agc.c:25-86
    float agc_ring[AGC_RB];
    int agc_out_index, agc_in_index;
    float agc_ring_max, agc_volts, agc_save_volts, agc_fast_backaverage, agc_hang_backaverage;
    int agc_hang_counter, agc_decay_type, agc_state;
    float agc_wold;
    float agc_sample_rate;   // for the "decimation rate changed" re-init rule, audio_agc.c:138-142
    int agc_action, agc_hang_action;
    // SAM
    float sam_fil_out, sam_lowpass, sam_omega2, sam_phs, sam_dsI, sam_dsQ;
    float sam_a[24], sam_b[24], sam_c[24], sam_d[24];
    int sam_count;
    float fade_dc27, fade_dc_insert;
    int carrier_freq_offset;
    // FM, audio_driver.c:1516-1531
    float fm_i_prev, fm_q_prev, fm_lpf_prev, fm_hpf_prev_a, fm_hpf_prev_b, fm_sql_avg;
    int fm_count, fm_squelched;
    float fm_gz[9], fm_subdet;               // Goertzel buf[3] of the HIGH / LOW / CTR subtone detectors, smoothed ratio (audio_driver.c:1665-1734)
    int fm_gcount, fm_tdet, fm_tone_detected;
    // LMS automatic notch: arm_lms_norm_f32 instance (coefficients, the newest 64 inputs as a circular window with
    // notch_head = slot of the oldest, energy, x0) and the decorrelation delay line of AudioDriver_NotchFilter
    float notch_coef[NOTCH_TAPS], notch_x[NOTCH_TAPS], notch_delay[NOTCH_DELAY];
    float notch_energy, notch_x0;
    int notch_head, notch_inbuf, notch_outbuf;
    // a_buffer[1] persistence is not needed: every non-FM path has an interpolator
    uint32_t samp_ptr;       // spectrum ring write pointer, audio_driver.c:1816-1824
    BiquadS zoom_bq_i[4], zoom_bq_q[4];   // IIR_biquad_Zoom_FFT_I/Q state (survives a reconfiguration)
    float zoom_hist_i[4], zoom_hist_q[4]; // DECIMATE_ZOOM_FFT_I/Q state (3 used; zeroed by AudioDriver_Spectrum_Set)
    int adc_clip, adc_half_clip, adc_quarter_clip;
    long long blocks;
    // AudioDriver_RxHandleTwinpeaks statics (audio_driver.c:2183-2187) and ts.twinpeaks_tested
    int tw_state, tw_counter, tw_runs, tw_restarts;
    float tw_phase;
};

// The sample-serial part of ChanState (same member names), the per-thread working copy of rx_serial.cu.
struct SerState {
    float interp_hist[INTERP_HIST];
    float pre_s[MAX_LAT], aa_s[MAX_LAT], sql_s[MAX_LAT];
    BiquadS bq1[4], bq2;
    float agc_ring[AGC_RB];
    int agc_out_index, agc_in_index;
    float agc_ring_max, agc_volts, agc_save_volts, agc_fast_backaverage, agc_hang_backaverage;
    int agc_hang_counter, agc_decay_type, agc_state;
    float agc_wold;
    int agc_action, agc_hang_action;
    float sam_fil_out, sam_lowpass, sam_omega2, sam_phs, sam_dsI, sam_dsQ;
    float sam_a[24], sam_b[24], sam_c[24], sam_d[24];
    int sam_count;
    float fade_dc27, fade_dc_insert;
    int carrier_freq_offset;
    float fm_i_prev, fm_q_prev, fm_lpf_prev, fm_hpf_prev_a, fm_hpf_prev_b, fm_sql_avg;
    int fm_count, fm_squelched;
    float fm_gz[9], fm_subdet;               // Goertzel buf[3] of the HIGH / LOW / CTR subtone detectors, smoothed ratio (audio_driver.c:1665-1734)
    int fm_gcount, fm_tdet, fm_tone_detected;
    // LMS automatic notch: arm_lms_norm_f32 instance (coefficients, the newest 64 inputs as a circular window with
    // notch_head = slot of the oldest, energy, x0) and the decorrelation delay line of AudioDriver_NotchFilter
    float notch_coef[NOTCH_TAPS], notch_x[NOTCH_TAPS], notch_delay[NOTCH_DELAY];
    float notch_energy, notch_x0;
    int notch_head, notch_inbuf, notch_outbuf;
};

// Spectral NR state, audio_nr.c (allocated only when a channel enables DSP_NR_ENABLE)
struct NrState {
    float bufs[4][256];      // mmb.nr_audio_buff[k]: [0..127] packed input, [128..255] processed output
    int trans_count_in, outbuff_count, fill_in_pt, out_buffer;
    int in_fifo[5], in_head, in_tail, out_fifo[5], out_head, out_tail;
    int current_buffer_idx, was_here;
    float dec_hist[4], int_hist[20];    // DECIMATE_NR (3 used) / INTERPOLATE_NR (19 used) histories
    float last_sample[128], last_ifft[128], Hk[128], Hk_old[128], Nest0[128], xt[128], pslp[128];
    int first_time, init_counter;
    float nb_work[128 + 26];  // alt_noise_blanking's static working_buffer (audio_nr.c:2282): 26 samples carried between frames
};

// TxProcessor_Run, SSB voice branch (tx_processor.c:891-1078): parameters and state
struct TxParams {
    int enabled;             // is_ssb(dmod_mode), or AM with a frequency-translate mode (tx_processor.c:996-1006)
    int am;                  // TxProcessor_AM: both sidebands + carrier after the Hilbert pair (:783-790)
    float alc_gain_scaling;  // SSB_ALC_GAIN_CORRECTION 1.00 / AM_ALC_GAIN_CORRECTION 0.23 (audio_driver.h:417, :428) / FM_ALC_GAIN_CORRECTION 0.95
    int fm;                  // TxProcessor_FM (tx_processor.c:534-589): pre-emphasis + NCO on the sine table instead of the Hilbert pair
    int fm_word;             // (65536 * |translate_freq|) / 48000, :569
    int fm_swap;             // translate_freq < 0: I and Q buffers swapped, :572-573
    float fm_mult;           // 2 for 5 kHz deviation, else 1
    int dds_off;             // pool offset of DDS_TABLE (1024 entries)
    uint32_t fm_sub_step, fm_burst_step;   // softdds steps of the sub-audible tone / the tone burst, 0 = off (tx_processor.c:554-564)
    float fm_sub_scale, fm_burst_scale;    // FM_SUBAUDIBLE_TONE_AMPLITUDE_SCALING / FM_TONE_BURST_AMPLITUDE_SCALING x deviation factor
    int lsb;                 // dmod_mode == DEMOD_LSB: I/Q filters swapped (tx_processor.c:477-478)
    float gain_calc;         // mic gain / MIC_GAIN_RESCALE * 2^-16 (tx_processor.c:360-381)
    LatticeP lat;            // IIR_TXFilter
    float bq[3][5];          // IIR_TX_biquad: treble shelf, bass shelf, pass-through
    int comp_enabled;        // ts.tx_comp_level > -1
    float postfilt_gain;     // alc_tx_postfilt_gain_var/2 + 0.5 (tx_processor.c:183)
    float alc_decay;         // ads.alc_decay (audio_management.c:15-19)
    int hil_ntaps, hil_ci, hil_cq;
    int shift_kind, shift_down, shift_freq;
    float osc_cos, osc_sin;
    float final_gain_i, final_gain_q;   // tx_processor.c:302-303 (x 65536 x SSB_GAIN_COMP)
    float phase_bal;         // ads.iq_phase_balance_tx[trans_idx]
};

struct TxState {
    float lat_s[MAX_LAT];
    BiquadS bq[3];
    float alc_val;
    float delay[320];        // audio_delay_buffer, AUDIO_DELAY_BUFSIZE = 5*IQ_BUFSZ
    uint32_t alc_delay_inbuf;
    float hist[H2];          // shared input history of the two 201-tap Hilbert filters
    float peak_audio;
    long long blocks;
    float fm_hpf_a, fm_hpf_b;   // TxProcessor_FM statics hpf_prev_a / hpf_prev_b / fm_mod_accum (tx_processor.c:537-538)
    uint32_t fm_accum;
    uint32_t fm_dds_sub_acc, fm_dds_burst_acc;   // soft_dds_t accumulators (reset by every configuration, softdds.c:38-45)
};

// Coefficients of the fused narrow-SSB kernel, passed by value as a kernel parameter so that
// every tap is an immediate constant-bank operand of its FFMA.
struct FusedCoefs {
    float dec[144];          // 83-tap sideband-suppression decimator (fir_rx_decimate_4.c:81) at [32, 115), zero padded
    float hil_i[224];        // 199-tap i_rx_new_coeffs (iq_rx_filter.c:589) at [12, 211), zero padded
    float hil_q[224];        // 199-tap q_rx_new_coeffs (iq_rx_filter.c:591) at [12, 211), zero padded
};

}  // namespace uhsdr
