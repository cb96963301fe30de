/*
 * ref_spectrum_harness.c -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Drives the reference's own spectrum-display state machine, UiSpectrum_RedrawSpectrum
 * (mchf-eclipse/drivers/ui/lcd/ui_spectrum.c:1350-1500), states 0-4: ring snapshot + Hann window, 512-point
 * FFT, magnitudes, IIR bin averaging (:1432-1446), dBm / dBm-per-Hz of the passband (UiSpectrum_CalculateDBm,
 * :1990-2122), log scaling with the sliding display offset (UiSpectrum_ScaleFFT :1258-1296,
 * UiSpectrum_ScaleFFT2SpectrumWidth :1300-1337, :1485).  ui_spectrum.c is #included from /root/reference
 * (via -I, see oracle/Makefile) so that its static functions are reachable; the LCD routines it draws with
 * are no-op stubs below -- drawing (state 5) is never reached.  Nothing of it is copied into this repository.
 */
#include "uhsdr_board.h"
#include "arm_math.h"
/* cmsis_gcc.h has been included by now; ui_spectrum.c:1364 calls __DSB(), an ARM barrier instruction */
#define __DSB() ((void)0)
#include "ui_spectrum.c"

#include "cw_decoder.h"
#include "psk.h"
#include "rtty.h"
#include "radio_management.h"

/* ---- objects and routines of the UI / LCD stack that ui_spectrum.c links against (never used by states 0-4) ---- */
disp_resolution_t disp_resolution = RESOLUTION_480_320;
DialFrequency df;
cw_config_t cw_decoder_config;
psk_ctrl_t psk_ctrl_config;
const psk_speed_item_t psk_speeds[PSK_SPEED_NUM];
rtty_ctrl_t rtty_ctrl_config;
const rtty_shift_item_t rtty_shifts[RTTY_SHIFT_NUM];
void UiMenu_MapColors(uint32_t color, char *options, volatile uint32_t *clr_ptr) { (void)color; (void)options; if (clr_ptr) *clr_ptr = 0; }
uint32_t RadioManagement_GetTXDialFrequency(void) { return 7100000u; }
uint32_t RadioManagement_GetRXDialFrequency(void) { return 7100000u; }
int32_t RadioManagement_GetCWDialOffset(void) { return 0; }
uint16_t UiLcdHy28_PrintText(uint16_t x, uint16_t y, const char *s, const uint32_t c, const uint32_t b, uchar f) { (void)x; (void)y; (void)s; (void)c; (void)b; (void)f; return 0; }
uint16_t UiLcdHy28_PrintTextRight(uint16_t x, uint16_t y, const char *s, const uint32_t c, const uint32_t b, uchar f) { (void)x; (void)y; (void)s; (void)c; (void)b; (void)f; return 0; }
uint16_t UiLcdHy28_PrintTextCentered(const uint16_t x, const uint16_t y, const uint16_t w, const char *t, uint32_t c, uint32_t b, uint8_t f) { (void)x; (void)y; (void)w; (void)t; (void)c; (void)b; (void)f; return 0; }
uint16_t UiLcdHy28_TextWidth(const char *s, uchar f) { (void)s; (void)f; return 0; }
uint16_t UiLcdHy28_TextHeight(uint8_t f) { (void)f; return 8; }
void UiLcdHy28_DrawStraightLine(ushort x, ushort y, ushort l, uchar d, ushort c) { (void)x; (void)y; (void)l; (void)d; (void)c; }
void UiLcdHy28_DrawStraightLineDouble(ushort x, ushort y, ushort l, uchar d, ushort c) { (void)x; (void)y; (void)l; (void)d; (void)c; }
void UiLcdHy28_DrawHorizLineWithGrad(ushort x, ushort y, ushort l, ushort g) { (void)x; (void)y; (void)l; (void)g; }
void UiLcdHy28_DrawFullRect(ushort x, ushort y, ushort h, ushort w, ushort c) { (void)x; (void)y; (void)h; (void)w; (void)c; }
void UiLcdHy28_BulkPixel_OpenWrite(ushort x, ushort w, ushort y, ushort h) { (void)x; (void)w; (void)y; (void)h; }
void UiLcdHy28_BulkPixel_CloseWrite(void) {}
void UiLcdHy28_BulkPixel_PutBuffer(uint16_t *p, uint32_t n) { (void)p; (void)n; }

/* UiSpectrum_InitSpectrumDisplayData (:955-1083) for the 480x320 layout with the settings the engine takes in
 * uhsdr_spectrum_display_cfg_t.  The ring write pointer and the magnification belong to the RX path
 * (audio_driver.c:1811-1909) and survive. */
int ref_spectrum_display_init(int spectrum_db_scale, int spectrum_agc_rate, int spectrum_filter, int dbm_constant, int scope_w)
{
    const uint32_t samp_ptr = sd.samp_ptr;
    const uint8_t magnify = sd.magnify;
    if (scope_w < 1 || scope_w > SPECTRUM_WIDTH_MAX) return -1;      /* sd.Old_PosData[SPECTRUM_WIDTH_MAX] is filled over scope.w, :1076-1079 */
    disp_resolution = RESOLUTION_480_320;
    ts.spectrum_db_scale = (uint8_t)spectrum_db_scale;
    ts.spectrum_agc_rate = (uint8_t)spectrum_agc_rate;
    ts.spectrum_filter = (uint8_t)spectrum_filter;
    ts.dbm_constant = dbm_constant;
    ts.txrx_mode = TRX_MODE_RX;
    ts.menu_mode = 0; ts.mem_disp = 0; ts.SpectrumResize_flag = 0; ts.VirtualKeysShown_flag = 0;
    slayout.scope.w = (uint16_t)scope_w;
    slayout.scope.h = 64; slayout.wfall.h = 64;
    UiSpectrum_InitSpectrumDisplayData();
    sd.samp_ptr = samp_ptr;
    sd.magnify = magnify;
    sd.display_offset = 0;
    for (int i = 0; i < 512; i++) sd.FFT_AVGData[i] = 0;
    return 0;
}

/* One pass of UiSpectrum_RedrawSpectrum through states 0..4 (:1362-1487) on the current ring contents.
 *   mags[512]  FFT_MagData after state 2            avg[512]   FFT_AVGData after state 3
 *   disp[scope_w] FFT_Samples after state 4          lvl[3]     sm.dbm_cur, sm.dbmhz_cur, sd.display_offset */
int ref_spectrum_redraw(float *mags, float *avg, float *disp, float *lvl)
{
    if (sd.fft_iq_len != 1024 || !sd.enabled) return -1;
    sd.state = 0;
    ts.dial_moved = 0;
    UiSpectrum_RedrawSpectrum();      /* 0: snapshot + window */
    UiSpectrum_RedrawSpectrum();      /* 1: FFT */
    UiSpectrum_RedrawSpectrum();      /* 2: magnitudes */
    if (mags) memcpy(mags, sd.FFT_MagData, 512 * sizeof(float));
    UiSpectrum_RedrawSpectrum();      /* 3: averaging + dBm */
    if (avg) memcpy(avg, sd.FFT_AVGData, 512 * sizeof(float));
    if (sd.state != 4) return -2;
    UiSpectrum_RedrawSpectrum();      /* 4: log scaling, width scaling, display offset */
    if (disp) memcpy(disp, sd.FFT_Samples, (size_t)slayout.scope.w * sizeof(float));
    if (lvl) { lvl[0] = sm.dbm_cur; lvl[1] = sm.dbmhz_cur; lvl[2] = sd.display_offset; }
    sd.state = 0;
    return 0;
}
