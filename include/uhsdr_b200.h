/*
 * uhsdr_b200.h -- C ABI of the B200-native batched UHSDR receiver/transmitter DSP engine.
 *
 * This is the drop-in boundary for UHSDR's audio/RF block path.  The reference has no plugin
 * or FFI layer: its boundary is three C functions plus global state:
 *
 *   void AudioDriver_I2SCallback(AudioSample_t *audio, IqSample_t *iq, AudioSample_t *audioDst,
 *                                int16_t blockSize);            mchf-eclipse/drivers/audio/audio_driver.h:651
 *   static void AudioDriver_RxProcessor(IqSample_t*, AudioSample_t*, uint16_t, bool external_mute)
 *                                                               mchf-eclipse/drivers/audio/audio_driver.c:2603
 *   void TxProcessor_Run(AudioSample_t*, IqSample_t*, AudioSample_t*, uint16_t, bool)
 *                                                               mchf-eclipse/drivers/audio/tx_processor.h:24
 *   void AudioDriver_SetProcessingChain(uint8_t dmod_mode, bool reset_dsp_nr)
 *                                                               mchf-eclipse/drivers/audio/audio_driver.c:1093
 *
 * The engine keeps that block contract (32 IqSample_t in -> 32 AudioSample_t out per channel and
 * block, 48 ksps, int32 words with the signal left-justified by 16 bits) but processes many
 * independent channels and many consecutive blocks per call.  All per-channel parameters the
 * reference reads from its globals (ts, ads, agc_wdsp_conf, nr_params, sd) travel in
 * uhsdr_chan_cfg_t.  The coefficient tables (FilterPathInfo[], filters/ *.c) are NOT embedded:
 * the host hands them over once as a table blob (uhsdr_tables.h), exactly as the firmware links
 * its own filters/ *.c.
 *
 * Plain C, plain pointers and sizes, integer return codes, no exceptions, no CPU fallback:
 * every entry point fails with UHSDR_ERR_CUDA / UHSDR_ERR_NO_DEVICE when no B200 is usable.
 */
#ifndef UHSDR_B200_H
#define UHSDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UHSDR_BLOCK_SIZE 32        /* IQ_BLOCK_SIZE, uhsdr_board_config.h:217 */
#define UHSDR_SAMPLE_RATE 48000    /* IQ_SAMPLE_RATE, uhsdr_board_config.h:207 */
#define UHSDR_NUM_FILTER_PATHS 87  /* AUDIO_FILTER_PATH_NUM, audio_filter.h:141 */
#define UHSDR_SPECTRUM_FFT_LEN 512 /* sd.fft_iq_len/2 on 480x320 displays, ui_spectrum.c:975-979 */

/* IqSample_t / AudioSample_t, audio_driver.h:44-52: two int32 words, l then r.
 * RX in:  l = I, r = Q (audio_driver.c:2677-2678).  RX out: l = main audio, r = copy
 * (audio_driver.c:2911-2912).  TX in: microphone in l (tx_processor.c:394).  TX out: l = I, r = Q. */
typedef struct { int32_t l; int32_t r; } uhsdr_iq_sample_t;
typedef struct { int32_t l; int32_t r; } uhsdr_audio_sample_t;

/* DemodModes_t, uhsdr_board.h:72-84 (same numeric values). */
enum {
    UHSDR_DEMOD_USB = 0, UHSDR_DEMOD_LSB = 1, UHSDR_DEMOD_CW = 2, UHSDR_DEMOD_AM = 3,
    UHSDR_DEMOD_SAM = 4, UHSDR_DEMOD_FM = 5, UHSDR_DEMOD_DIGI = 6
};
/* sam_sideband_t, audio_driver.h:181-190. */
enum { UHSDR_SAM_SIDEBAND_BOTH = 0, UHSDR_SAM_SIDEBAND_LSB = 1, UHSDR_SAM_SIDEBAND_USB = 2 };
/* ts.iq_freq_mode, audio_driver.h:520-526. */
enum {
    UHSDR_FREQ_IQ_CONV_OFF = 0, UHSDR_FREQ_IQ_CONV_P6KHZ = 1, UHSDR_FREQ_IQ_CONV_M6KHZ = 2,
    UHSDR_FREQ_IQ_CONV_P12KHZ = 3, UHSDR_FREQ_IQ_CONV_M12KHZ = 4
};
/* ts.dsp.active bits, audio_driver.h:195-200. */
enum {
    UHSDR_DSP_NR_ENABLE = 0x01, UHSDR_DSP_NR_POSTAGC_ENABLE = 0x02, UHSDR_DSP_NOTCH_ENABLE = 0x04,
    UHSDR_DSP_NB_ENABLE = 0x08, UHSDR_DSP_MNOTCH_ENABLE = 0x10, UHSDR_DSP_MPEAK_ENABLE = 0x20
};
/* ts.tx_filter, tx_processor.c:92-102. */
enum { UHSDR_TX_FILTER_SOPRANO = 1, UHSDR_TX_FILTER_TENOR = 2, UHSDR_TX_FILTER_BASS = 3 };

/* Return codes. */
enum {
    UHSDR_OK = 0,
    UHSDR_ERR_ARG = -1,        /* NULL pointer, channel/path out of range, nblocks <= 0 ...       */
    UHSDR_ERR_NO_DEVICE = -2,  /* no CUDA device / not an sm_100 part                              */
    UHSDR_ERR_CUDA = -3,       /* a CUDA runtime call failed; see uhsdr_last_error()               */
    UHSDR_ERR_TABLES = -4,     /* table blob missing, wrong magic/version or truncated             */
    UHSDR_ERR_UNSUPPORTED = -5,/* configuration the engine does not implement (fails loudly)       */
    UHSDR_ERR_STATE = -6       /* channel not configured                                           */
};

/* Per-channel configuration: exactly the fields the reference's block path reads from its globals.
 * Defaults (uhsdr_default_chan_cfg) are the reference's ui_configuration.c:60-220 values. */
typedef struct {
    uint32_t struct_size;        /* sizeof(uhsdr_chan_cfg_t), for ABI evolution                     */
    /* mode / path selection */
    int32_t dmod_mode;           /* ts.dmod_mode                                                    */
    int32_t filter_path;         /* ts.filter_path: index into FilterPathInfo[87], audio_filter.c:147 */
    int32_t cw_lsb;              /* ts.cw_lsb, radio_management.c:1676                              */
    int32_t digi_lsb;            /* ts.digi_lsb                                                     */
    int32_t iq_freq_mode;        /* ts.iq_freq_mode, default FREQ_IQ_CONV_M12KHZ                    */
    /* IQ correction, audio_driver.c:2254-2316 */
    int32_t iq_auto_correction;  /* ts.iq_auto_correction (default 1)                               */
    float   rx_adj_gain_i;       /* ts.rx_adj_gain_var.i (manual mode)                              */
    float   rx_adj_gain_q;       /* ts.rx_adj_gain_var.q                                            */
    float   iq_phase_balance_rx; /* ads.iq_phase_balance_rx                                         */
    /* DSP switches and EQ, audio_driver.c:994-1050 */
    int32_t dsp_active;          /* ts.dsp.active bit mask                                          */
    int32_t notch_frequency;     /* ts.dsp.notch_frequency (Hz)                                     */
    int32_t peak_frequency;      /* ts.dsp.peak_frequency (Hz)                                      */
    int32_t bass_gain;           /* ts.dsp.bass_gain (dB), default 2                                */
    int32_t treble_gain;         /* ts.dsp.treble_gain (dB), default 0                              */
    int32_t nr_strength;         /* ts.dsp.nr_strength, default 160                                 */
    int32_t nb_setting;          /* ts.dsp.nb_setting                                               */
    /* WDSP AGC, agc_wdsp_params_t audio_agc.h:23-36 */
    int32_t agc_mode;            /* 0..5 (5 = off), default 2                                       */
    int32_t agc_slope;           /* default 70                                                      */
    int32_t agc_hang_enable;     /* default 0                                                       */
    int32_t agc_thresh;          /* default 20                                                      */
    int32_t agc_hang_thresh;     /* 45, audio_agc.c:113                                             */
    int32_t agc_hang_time;       /* 500, audio_agc.c:112                                            */
    int32_t agc_tau_decay[6];    /* {4000,2000,500,250,50,1}                                        */
    int32_t agc_tau_hang_decay;  /* 500                                                             */
    /* SAM, audio_driver.c:709-745 */
    int32_t sam_sideband;        /* ads.sam_sideband                                                */
    int32_t sam_fade_leveler;    /* ads.fade_leveler, default 1                                     */
    int32_t sam_pll_fmax;        /* ads.pll_fmax_int, default 2500                                  */
    int32_t sam_zeta;            /* ads.zeta_int (x100), default 65                                 */
    int32_t sam_omegaN;          /* ads.omegaN_int, default 250                                     */
    /* FM, audio_driver.c:1544-1737 */
    int32_t fm_sql_threshold;    /* ts.fm_sql_threshold, default 12                                 */
    int32_t fm_dev_5khz;         /* ts.flags2 & FLAGS2_FM_MODE_DEVIATION_5KHZ                       */
    float   fm_subaudible_tone_det_freq; /* ads.fm_conf.subaudible_tone_det_freq, 0 = off           */
    /* spectral NR, audio_nr.c:1841-2195 */
    int32_t nr_decimation_enable;/* nr_params.NR_decimation_enable (default 1)                      */
    /* spectrum display, ui_spectrum.c:1350-1390 */
    int32_t spectrum_enable;     /* 0 = no spectrum ring / FFT for this channel                     */
    int32_t spectrum_magnify;    /* sd.magnify 0..5 (1..5 = zoom FFT)                               */
    float   codec_gain_calc;     /* ads.codec_gain_calc (spectrum scaling), default 1               */
    /* TX, tx_processor.c */
    int32_t tx_filter;           /* ts.tx_filter, default SOPRANO                                   */
    int32_t tx_bass_gain;        /* ts.dsp.tx_bass_gain, default 4                                  */
    int32_t tx_treble_gain;      /* ts.dsp.tx_treble_gain, default 4                                */
    int32_t tx_mic_gain;         /* ts.tx_gain[TX_AUDIO_MIC], default 15                            */
    int32_t tx_comp_level;       /* ts.tx_comp_level, default 2                                     */
    int32_t tx_alc_decay;        /* ts.alc_decay, default 10                                        */
    int32_t tx_alc_postfilt_gain;/* ts.alc_tx_postfilt_gain, default 1                              */
    float   tx_power_factor;     /* ts.tx_power_factor                                              */
    float   tx_adj_gain_i;       /* ts.tx_adj_gain_var[trans_idx].i                                 */
    float   tx_adj_gain_q;       /* ts.tx_adj_gain_var[trans_idx].q                                 */
    float   iq_phase_balance_tx; /* ads.iq_phase_balance_tx[trans_idx]                              */
    /* LMS automatic notch (DSP_NOTCH_ENABLE), audio_driver.c:1165-1187, :1746-1763.  The tap count and the
     * decorrelation delay are fixed by the firmware (DSP_NOTCH_NUMTAPS_MIN == MAX == 64,
     * DSP_NOTCH_BUFLEN_MIN == MAX == 128, audio_driver.h:486-492); only the convergence rate is a setting. */
    int32_t notch_mu;            /* ts.dsp.notch_mu, 0..40, default 10 (DSP_NOTCH_MU_DEFAULT)       */
    /* FM transmit tones, tx_processor.c:554-564 (softdds single tones added to the pre-emphasised audio) */
    float   fm_subaudible_tone_gen_freq; /* ads.fm_conf.subaudible_tone_gen_freq in Hz, 0 = off (default)  */
    int32_t fm_tone_burst_mode;  /* ts.fm_tone_burst_mode while the burst is keyed: 0 off, 1 = 1750 Hz, 2 = 2135 Hz */
} uhsdr_chan_cfg_t;

/* Per-channel side outputs (SURVEY.md section 5 "metrics"); none is on the parity-critical path. */
typedef struct {
    int32_t adc_clip;            /* ads.adc_clip / half / quarter, audio_driver.c:2662-2675         */
    int32_t adc_half_clip;
    int32_t adc_quarter_clip;
    int32_t agc_action;          /* agc_wdsp_conf.action, audio_agc.c:552-561                       */
    int32_t agc_hang_action;     /* agc_wdsp_conf.hang_action, audio_agc.c:400-407                  */
    int32_t fm_squelched;        /* ads.fm_conf.squelched                                           */
    float   fm_sql_avg;          /* ads.fm_conf.sql_avg                                             */
    int32_t sam_carrier_freq_offset; /* ads.carrier_freq_offset, audio_driver.c:2150-2162           */
    float   iq_corr_c1;          /* adb.iq_corr.M_c1, M_c2                                          */
    float   iq_corr_c2;
    float   tx_peak_audio;       /* ads.peak_audio, tx_processor.c:403                              */
    float   tx_alc_val;          /* ads.alc_val                                                     */
    int64_t blocks_processed;
    /* "twin peaks" detector of the automatic IQ correction (AudioDriver_RxHandleTwinpeaks, audio_driver.c:2173-2248):
     * ts.twinpeaks_tested -- 2 = waiting for the channel to settle (1000 blocks), 0 = sampling the I/Q phase error (50 blocks),
     * 1 = done, phase error below 22.5 degrees, 4 = codec restart requested (phase error above), 3 = uncorrectable (4th request).
     * The firmware's main loop answers 4 by restarting the codec and re-arming the detector (ui_driver.c:7422-7425):
     * uhsdr_twinpeaks_rearm is that re-arm. */
    int32_t twinpeaks_state;
    int32_t twinpeaks_restarts;  /* codec_restarts, audio_driver.c:2186 */
} uhsdr_chan_status_t;
enum { UHSDR_TWINPEAKS_SAMPLING = 0, UHSDR_TWINPEAKS_DONE = 1, UHSDR_TWINPEAKS_WAIT = 2, UHSDR_TWINPEAKS_UNCORRECTABLE = 3, UHSDR_TWINPEAKS_CODEC_RESTART = 4 };

typedef struct uhsdr_engine uhsdr_engine_t;

/* Library identity.  uhsdr_b200_backend() returns "cuda-sm100a". */
int         uhsdr_b200_abi_version(void);   /* 2: + tables_validate, spectrum_display, twinpeaks, multi */
const char *uhsdr_b200_backend(void);
const char *uhsdr_strerror(int code);
/* Text of the last CUDA failure on this engine (or of engine creation when e == NULL). */
const char *uhsdr_last_error(const uhsdr_engine_t *e);

/* Reference defaults (ui_configuration.c:60-220; USB, path 35, -12 kHz translate, AGC mode 2). */
int uhsdr_default_chan_cfg(uhsdr_chan_cfg_t *cfg);

/* Checks a table blob (uhsdr_tables.h) without touching a device: magic / version / size, every section and
 * every cross-index inside the blob.  UHSDR_OK or UHSDR_ERR_TABLES (text in uhsdr_last_error(NULL)).
 * uhsdr_engine_create runs the same check. */
int uhsdr_tables_validate(const void *tables, size_t tables_bytes);

/* Engine for num_channels channels on CUDA device `device`.  `tables` is the blob described in
 * uhsdr_tables.h (the reference's FilterPathInfo[] + coefficient arrays); it is copied.
 * Replaces AudioDriver_Init (audio_driver.c:677). */
int uhsdr_engine_create(uhsdr_engine_t **out, int num_channels, int device,
                        const void *tables, size_t tables_bytes);
int uhsdr_engine_destroy(uhsdr_engine_t *e);
int uhsdr_engine_num_channels(const uhsdr_engine_t *e);

/* AudioDriver_SetProcessingChain (audio_driver.c:1093) + TxProcessor_Set (tx_processor.c:72) for
 * channels [first, first+count).  reset != 0: fresh-process semantics (all DSP state zeroed to the
 * boot values, SURVEY.md 8a "state inventory"); reset == 0: the reference's reconfigure semantics
 * (FIR/decimator/interpolator histories and lattice states cleared, the rest kept). */
int uhsdr_configure_channels(uhsdr_engine_t *e, int first, int count,
                             const uhsdr_chan_cfg_t *cfg, int reset);
int uhsdr_configure_channel(uhsdr_engine_t *e, int channel, const uhsdr_chan_cfg_t *cfg, int reset);
/* Same for channels first, first+stride, first+2*stride, ... (count of them): one call configures
 * e.g. every even channel of an interleaved USB / LSB plan. */
int uhsdr_configure_channels_strided(uhsdr_engine_t *e, int first, int count, int stride,
                                     const uhsdr_chan_cfg_t *cfg, int reset);

/* AudioDriver_RxProcessor for every channel and nblocks consecutive blocks.
 * iq / audio: channel-major, [num_channels][nblocks*32] elements, HOST memory (pinned or pageable).
 * mute: optional [num_channels][nblocks] bytes = external_mute per block (NULL = never muted).
 * Copies in, runs the CUDA chain, copies out, returns when the audio is in `audio`. */
int uhsdr_rx_process(uhsdr_engine_t *e, const uhsdr_iq_sample_t *iq, uhsdr_audio_sample_t *audio,
                     int nblocks, const uint8_t *mute);
/* Same with DEVICE pointers on the engine's device; asynchronous on the engine's stream
 * (uhsdr_engine_sync to wait).  audio_f (optional, device) receives the float audio before the
 * int32 formatting, [num_channels][nblocks*32] floats (adb.a_buffer[1], audio_driver.c:2911).
 * Alignment: any 8-byte aligned iq_dev / audio_dev and 4-byte aligned audio_f_dev are accepted; the fast kernels need
 * iq_dev and audio_dev 32-byte aligned and audio_f_dev 16-byte aligned (cudaMalloc gives 256), other alignments take
 * the slower general kernels with identical results within the build's tolerance. */
int uhsdr_rx_process_device(uhsdr_engine_t *e, const uhsdr_iq_sample_t *iq_dev,
                            uhsdr_audio_sample_t *audio_dev, float *audio_f_dev,
                            int nblocks, const uint8_t *mute_dev);
/* TxProcessor_Run, SSB voice branch (tx_processor.c:891, :989-995), AM branch (:996-1006) and FM branch (:1007-1016; AM and
 * FM need a frequency-translate mode): mic audio in, I/Q out.  Other modes (CW, digital): UHSDR_ERR_UNSUPPORTED. */
int uhsdr_tx_process(uhsdr_engine_t *e, const uhsdr_audio_sample_t *audio, uhsdr_iq_sample_t *iq,
                     int nblocks, const uint8_t *mute);
int uhsdr_tx_process_device(uhsdr_engine_t *e, const uhsdr_audio_sample_t *audio_dev,
                            uhsdr_iq_sample_t *iq_dev, float *iq_f_dev, int nblocks,
                            const uint8_t *mute_dev);
int uhsdr_engine_sync(uhsdr_engine_t *e);
/* The engine's CUDA stream as a void* (cudaStream_t), for callers that time with CUDA events. */
void *uhsdr_engine_stream(uhsdr_engine_t *e);

/* UiSpectrum_RedrawSpectrum states 0-2 (ui_spectrum.c:1362-1390): snapshot of the channel's
 * spectrum ring, Hann window, 512-point complex FFT, magnitudes.  mags: [count][512] floats (host). */
int uhsdr_get_spectrum(uhsdr_engine_t *e, int first, int count, float *mags);
int uhsdr_get_spectrum_device(uhsdr_engine_t *e, int first, int count, float *mags_dev);

/* UiSpectrum_RedrawSpectrum states 0-4 (ui_spectrum.c:1362-1487), the display-side post-processing of the spectrum:
 * the FFT above, IIR averaging of the bins (:1432-1446), log scaling with the sliding display offset in frequency order
 * (UiSpectrum_ScaleFFT :1258-1296, :1485), rescaling to the scope width (UiSpectrum_ScaleFFT2SpectrumWidth :1300-1337), and
 * the S-meter basis UiSpectrum_CalculateDBm (:1990-2122): dBm and dBm/Hz of the signal inside the filter passband.
 * The settings are the reference's own (ts.spectrum_db_scale / spectrum_agc_rate / spectrum_filter / dbm_constant,
 * ui_configuration.c:94,137,138,206; slayout.scope.w).  The averaged bins and the display offset persist per channel. */
typedef struct {
    int32_t struct_size;
    int32_t spectrum_db_scale;   /* 1..8 = 5 / 7.5 / 10 / 15 / 20 dB, 1 / 2 / 3 S-units per division; default 3 (DB_DIV_10)   */
    int32_t spectrum_agc_rate;   /* 1..50, default 25 (SPECTRUM_SCOPE_AGC_DEFAULT)                                            */
    int32_t spectrum_filter;     /* 1..20, default 4 (SPECTRUM_FILTER_DEFAULT)                                                */
    int32_t dbm_constant;        /* -100..100, default 0                                                                      */
    int32_t scope_width;         /* slayout.scope.w: display columns, 1..512 (reference: <= 480, default 480); 512 = one per bin */
} uhsdr_spectrum_display_cfg_t;
typedef struct {
    float dbm;                   /* sm.dbm_cur,   ui_spectrum.c:2114 */
    float dbmhz;                 /* sm.dbmhz_cur, ui_spectrum.c:2115 */
    float display_offset;        /* sd.display_offset after the update of :1485 */
} uhsdr_spectrum_level_t;
int uhsdr_default_spectrum_display_cfg(uhsdr_spectrum_display_cfg_t *cfg);
/* disp: [count][scope_width] floats (sd.FFT_Samples after state 4); levels: [count]; avg (optional): [count][512] floats
 * (sd.FFT_AVGData after state 3).  Host buffers. */
int uhsdr_spectrum_display(uhsdr_engine_t *e, int first, int count, const uhsdr_spectrum_display_cfg_t *cfg,
                           float *disp, uhsdr_spectrum_level_t *levels, float *avg);
/* Same with device buffers, asynchronous on the engine's stream; mags_dev (optional): [count][512] magnitudes of state 2. */
int uhsdr_spectrum_display_device(uhsdr_engine_t *e, int first, int count, const uhsdr_spectrum_display_cfg_t *cfg,
                                  float *disp_dev, uhsdr_spectrum_level_t *levels_dev, float *avg_dev, float *mags_dev);

int uhsdr_get_status(uhsdr_engine_t *e, int first, int count, uhsdr_chan_status_t *status);
/* ts.twinpeaks_tested = TWINPEAKS_WAIT for channels [first, first+count) (after the host has "restarted the codec"). */
int uhsdr_twinpeaks_rearm(uhsdr_engine_t *e, int first, int count);

/* ---- several GPUs of one box behind one handle -------------------------------------------------------
 * Channels are independent, so global channel c lives on exactly one device: device g of G owns the contiguous range
 * uhsdr_channel_range(g, G, num_channels).  There is no collective and no peer traffic; a call is forwarded to the per-device
 * engines from one host thread per device.  Buffers are host memory, channel-major over ALL channels. */
typedef struct uhsdr_multi uhsdr_multi_t;
/* [first, first+count) of `total` channels owned by `rank` of `world`; sizes differ by at most one. */
int uhsdr_channel_range(int rank, int world, int total, int *first, int *count);
int uhsdr_multi_create(uhsdr_multi_t **out, int num_channels, const int *devices, int num_devices,
                       const void *tables, size_t tables_bytes);
int uhsdr_multi_destroy(uhsdr_multi_t *m);
int uhsdr_multi_num_devices(const uhsdr_multi_t *m);
int uhsdr_multi_num_channels(const uhsdr_multi_t *m);
const char *uhsdr_multi_last_error(const uhsdr_multi_t *m);
/* The engine of device slot `index` and the global channel range it owns (for the device-pointer entry points). */
uhsdr_engine_t *uhsdr_multi_engine(uhsdr_multi_t *m, int index, int *first, int *count);
/* Global channel numbers; forwarded to the owning engines. */
int uhsdr_multi_configure_channels(uhsdr_multi_t *m, int first, int count, const uhsdr_chan_cfg_t *cfg, int reset);
int uhsdr_multi_configure_channels_strided(uhsdr_multi_t *m, int first, int count, int stride,
                                           const uhsdr_chan_cfg_t *cfg, int reset);
int uhsdr_multi_rx_process(uhsdr_multi_t *m, const uhsdr_iq_sample_t *iq, uhsdr_audio_sample_t *audio,
                           int nblocks, const uint8_t *mute);
int uhsdr_multi_tx_process(uhsdr_multi_t *m, const uhsdr_audio_sample_t *audio, uhsdr_iq_sample_t *iq,
                           int nblocks, const uint8_t *mute);
int uhsdr_multi_get_status(uhsdr_multi_t *m, int first, int count, uhsdr_chan_status_t *status);

/* Number of kernel launches issued by this engine so far (bench.py "gpu_launches"). */
int64_t uhsdr_engine_launch_count(const uhsdr_engine_t *e);

#ifdef __cplusplus
}
#endif
#endif /* UHSDR_B200_H */
