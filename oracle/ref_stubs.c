/*
 * ref_stubs.c -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * The stub surface needed to link the reference's own hot-path sources on x86-64
 * (SURVEY.md section 8c): globals the firmware defines in UI/board files that are out of scope,
 * the four mode predicates restated from the reference, and no-op stand-ins for the
 * CW/RTTY/PSK/FreeDV/USB/LED side consumers.
 */
#include "uhsdr_board.h"
#include "ui_driver.h"
#include "profiling.h"
#include "audio_driver.h"
#include "audio_nr.h"
#include "radio_management.h"
#include "ui_spectrum.h"
#include "freedv_uhsdr.h"
#include "cw_gen.h"
#include "cw_decoder.h"
#include "rtty.h"
#include "psk.h"
#include "usbd_audio_if.h"
#include "uhsdr_hw_i2s.h"

/* globals owned by out-of-scope translation units */
__IO TransceiverState ts;              /* hardware/uhsdr_board.c */
SpectrumDisplay sd;                    /* drivers/ui/lcd/ui_spectrum.c */
MultiModeBuffer_t mmb;                 /* drivers/audio/freedv_uhsdr.c */
EventProfile_t eventProfile;           /* misc/profiling.c */
freedv_conf_t freedv_conf;
oracle_scb_t shim_scb;

/* RingBuffer_Declare() only declares; give the three FreeDV ring buffers storage. */
RingBuffer_data_t fdv_iq_rb, fdv_audio_rb, fdv_demod_rb;

/* radio_management.c:1643-1664 */
bool RadioManagement_UsesBothSidebands(uint16_t dmod_mode)
{
    bool r = (dmod_mode == DEMOD_AM) || (dmod_mode == DEMOD_SAM && ads.sam_sideband == SAM_SIDEBAND_BOTH)
             || (dmod_mode == DEMOD_FM);
    r = r || (dmod_mode == DEMOD_SSBSTEREO) || (dmod_mode == DEMOD_IQ)
          || (dmod_mode == DEMOD_SAM && ads.sam_sideband == SAM_SIDEBAND_STEREO);
    return r;
}
/* radio_management.c:1666-1690 */
bool RadioManagement_LSBActive(uint16_t dmod_mode)
{
    switch (dmod_mode) {
    case DEMOD_SAM:  return ads.sam_sideband == SAM_SIDEBAND_LSB;
    case DEMOD_LSB:  return true;
    case DEMOD_CW:   return ts.cw_lsb;
    case DEMOD_DIGI: return ts.digi_lsb;
    default:         return false;
    }
}
/* radio_management.c:1962 */
bool RadioManagement_FmDevIs5khz(void) { return (ts.flags2 & FLAGS2_FM_MODE_DEVIATION_5KHZ) != 0; }
/* radio_management.c:587: CW and the text modems transmit at zero IF; the oracle drives voice SSB only. */
bool RadioManagement_IsTxAtZeroIF(uint8_t dmod_mode, uint8_t digital_mode) { (void)digital_mode; return dmod_mode == DEMOD_CW; }
bool RadioManagement_UsesTxSidetone(void) { return ts.dmod_mode == DEMOD_CW; }

/* ui_driver.c:395-433 */
bool is_dsp_nr(void) { return (ts.dsp.active & DSP_NR_ENABLE) != 0; }
bool is_dsp_nb_active(void) { return ((ts.dsp.active & DSP_NB_ENABLE) != 0) && (ts.dsp.nb_setting > 0); }
bool is_dsp_mnotch(void) { return (ts.dsp.active & DSP_MNOTCH_ENABLE) != 0; }
bool is_dsp_mpeak(void) { return (ts.dsp.active & DSP_MPEAK_ENABLE) != 0; }

/* side consumers / hardware: no-ops */
void UsbdAudio_PutSample(int16_t sample) { (void)sample; }
void UsbdAudio_FillTxBuffer(AudioSample_t *buffer, uint32_t len) { (void)buffer; (void)len; }
void Board_GreenLed(ledstate_t state) { (void)state; }
void UiDriver_Callback_AudioISR(void) {}
void UhsdrHwI2s_Codec_ClearTxDmaBuffer(void) {}
void CwGen_Init(void) {}
bool CwGen_Process(float32_t *i, float32_t *q, uint32_t size) { (void)i; (void)q; (void)size; return false; }
void CwDecode_RxProcessor(float32_t *const src, int16_t blockSize) { (void)src; (void)blockSize; }
void CwDecode_Filter_Set(void) {}
void Rtty_Modem_Init(uint32_t r) { (void)r; }
void Rtty_Demodulator_ProcessSample(float32_t s) { (void)s; }
int16_t Rtty_Modulator_GenSample(void) { return 0; }
void Psk_Modem_Init(uint32_t r) { (void)r; }
void Psk_Demodulator_ProcessSample(float32_t s) { (void)s; }
int16_t Psk_Modulator_GenSample(void) { return 0; }
int32_t FreeDV_Iq_Get_FrameLen(void) { return 0; }
int32_t RingBuffer_GetData(RingBuffer_data_t *buf) { (void)buf; return 0; }
bool RingBuffer_PutSamples(RingBuffer_data_t *buf, void *samples, int32_t len) { (void)buf; (void)samples; (void)len; return false; }
bool RingBuffer_GetSamples(RingBuffer_data_t *buf, void *samples, int32_t len) { (void)buf; (void)samples; (void)len; return false; }

/* newlib's pow10f; glibc's former pow10f was an alias of exp10f. */
float exp10f(float);
float pow10f(float x) { return exp10f(x); }

/* CMSIS ships arm_bitreversal_32 only as Thumb assembly
 * (DSP_Lib/Source/TransformFunctions/arm_bitreversal2.S:86-110): for each table pair (a,b) of BYTE
 * offsets, swap the two 32-bit words at a, a+4 with those at b, b+4. */
void arm_bitreversal_32(uint32_t *pSrc, const uint16_t bitRevLen, const uint16_t *pBitRevTab)
{
    uint32_t n = ((uint32_t)bitRevLen + 1u) >> 1;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t *pa = (uint32_t *)((uint8_t *)pSrc + pBitRevTab[2 * i]);
        uint32_t *pb = (uint32_t *)((uint8_t *)pSrc + pBitRevTab[2 * i + 1]);
        uint32_t t0 = pa[0], t1 = pa[1];
        pa[0] = pb[0]; pa[1] = pb[1];
        pb[0] = t0; pb[1] = t1;
    }
}
