/*
 * uhsdr_port.h -- ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement ("port") of the reference's RX/TX block path with explicit per-channel
 * state, so that many channels can be run in one process and the oracle can travel to machines
 * where /root/reference does not exist.  It is pinned against the reference's own object code
 * (oracle/_ref/libuhsdr_ref.so) by tests/test_oracle_pin.py and against the committed golden
 * vectors in tests/golden/.  Nothing under uhsdr_b200/ may include or link this.
 */
#ifndef UHSDR_PORT_H
#define UHSDR_PORT_H

#include <stddef.h>
#include <stdint.h>
#include "uhsdr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct port_tables port_tables_t;
typedef struct port_chan port_chan_t;

port_tables_t *port_tables_load(const void *blob, size_t bytes);
void port_tables_free(port_tables_t *t);

/* Fresh channel (firmware boot + AudioDriver_SetProcessingChain). NULL on unsupported cfg. */
port_chan_t *port_chan_create(const port_tables_t *t, const uhsdr_chan_cfg_t *cfg);
/* AudioDriver_SetProcessingChain on a live channel (reference reconfigure semantics). */
int port_chan_reconfigure(port_chan_t *c, const uhsdr_chan_cfg_t *cfg);
void port_chan_free(port_chan_t *c);

/* nblocks x AudioDriver_RxProcessor. iq/audio: nblocks*32 x {int32 l, int32 r}; audio_f optional. */
int port_rx(port_chan_t *c, const int32_t *iq, int32_t *audio, float *audio_f, int nblocks,
            const uint8_t *mute);
/* nblocks x TxProcessor_Run (SSB voice). mic/iq: nblocks*32 x {l, r}; iq_f optional [n][2]. */
int port_tx(port_chan_t *c, const int32_t *mic, int32_t *iq, float *iq_f, int nblocks,
            const uint8_t *mute);
/* UiSpectrum_RedrawSpectrum states 0-2: 512 magnitudes. */
int port_spectrum(port_chan_t *c, float *mags);
/* UiSpectrum_RedrawSpectrum states 0-4: mags_out / avg_out optional [512]; disp [scope_width]; lvl [3] = dBm, dBm/Hz, display offset. */
int port_spectrum_display(port_chan_t *c, const uhsdr_spectrum_display_cfg_t *dc, float *mags_out, float *avg_out, float *disp, float *lvl);
int port_twinpeaks_rearm(port_chan_t *c);
int port_get_status(const port_chan_t *c, uhsdr_chan_status_t *st);

#ifdef __cplusplus
}
#endif
#endif
