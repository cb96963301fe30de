// kernels.h -- launch interfaces of the CUDA kernels (host side sees only these).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "uhsdr_dev.h"

namespace uhsdr {

struct RxArgs {
    const ChanParams *params;   // [num_channels]
    ChanState *state;           // [num_channels]
    NrState *nr;                // [num_channels] or nullptr
    float *spec_ring;           // [num_channels][1024] or nullptr
    const float *pool;          // coefficient pool
    const void *iq;             // [num_channels][nblocks*32] {int32 l, int32 r}
    void *audio;                // same shape
    float *audio_f;             // optional [num_channels][nblocks*32]
    const uint8_t *mute;        // optional [num_channels][nblocks]
    const int *chan_list;       // optional: channels handled by this launch (nullptr = 0..num_items-1)
    int num_items;
    int nblocks;
    long long chan_stride;      // samples between consecutive channels in iq / audio / audio_f (>= nblocks*32)
    long long mute_stride;      // bytes between consecutive channels in mute (>= nblocks)
    float *scratch;             // split path: FIR-stage outputs, [num_items][scratch_stride] floats
    long long scratch_stride;
    // tensor-core kernel only:
    float2 *iqc_log;            // optional [num_items][16]: (M_c1, M_c2) of the last min(16, nblocks) blocks, for the spectrum tap kernel
    int nr_handoff;             // 1: stop at the AGC output and write it to scratch [num_items][nblocks*8] (spectral NR + serial phase 2 follow)
};

cudaError_t launch_rx_generic(const RxArgs &a, cudaStream_t stream);
// split general path: time-parallel front end + FIR stages (one warp per channel, rx_generic.cu) into
// a.scratch, then the sample-serial stages with one channel per thread (rx_serial.cu)
cudaError_t launch_rx_front(const RxArgs &a, cudaStream_t stream);
// second-generation front kernel (rx_front2.cu): register-blocked FIRs, 512-sample chunks; same contract as launch_rx_front
bool rx_front2_eligible(const ChanParams &p);
cudaError_t launch_rx_front2(const RxArgs &a, cudaStream_t stream);
cudaError_t launch_rx_serial(const RxArgs &a, int phase, cudaStream_t stream);   // phase: see rx_serial.cu
// second-generation serial kernel (rx_serial2.cu, shipping build): same contract and phases
bool rx_serial2_eligible(const ChanParams &p);
cudaError_t launch_rx_serial2(const RxArgs &a, int phase, cudaStream_t stream);
cudaError_t launch_rx_nr(const RxArgs &a, cudaStream_t stream);                   // spectral NR on a.scratch, in place
int rx_split_floats_per_block(const ChanParams &p);     // scratch floats per 32-sample block and channel

// fused narrow-SSB receiver (rx_ssb_fused.cu)
bool fused_eligible(const ChanParams &p);
void fill_fused_coefs(FusedCoefs *fc, const float *dec83, const float *hil_i199, const float *hil_q199);
cudaError_t launch_rx_ssb_fused(const RxArgs &a, const FusedCoefs &fc, int sm_count, cudaStream_t stream);
// same chain with the 83-tap decimator and the 199-tap Hilbert pair on the tensor cores (rx_ssb_tc.cu; shipping build
// only).  dec_c / hil_ci / hil_cq: pool offsets of the decimator taps and of the two Hilbert tap sets.
bool rx_ssb_tc_available();
cudaError_t launch_rx_ssb_tc(const RxArgs &a, int dec_c, int hil_ci, int hil_cq, int sm_count, cudaStream_t stream);

struct TxArgs {
    const ChanParams *params;
    ChanState *state;           // shares the FreqShift NCO with RX (freq_shift.c:277-283 statics)
    TxState *tx;
    const TxParams *txp;
    const float *pool;
    const void *audio;          // [num_channels][nblocks*32] {int32 l, int32 r}, microphone in l
    void *iq;                   // [num_channels][nblocks*32] {int32 I, int32 Q}
    float *iq_f;                // optional [num_channels][nblocks*32][2]
    const uint8_t *mute;
    int num_items;
    int nblocks;
    float *scratch;             // split path: output of the serial stages, [num_channels][nblocks*32] floats (nullptr = single kernel)
    long long chan_stride;      // samples between consecutive channels in audio / iq / iq_f (>= nblocks*32)
    long long mute_stride;      // bytes between consecutive channels in mute
};
// scratch == nullptr: the whole modulator in one kernel.  Otherwise launch_tx_serial (one channel per thread:
// AudioBufferFill, lattice, biquads, compressor -> scratch) first, then launch_tx_ssb (Hilbert pair, translation, output).
cudaError_t launch_tx_ssb(const TxArgs &a, cudaStream_t stream);
cudaError_t launch_tx_serial(const TxArgs &a, cudaStream_t stream);
cudaError_t launch_tx_boot(TxState *tx, int n, cudaStream_t stream);
cudaError_t launch_nr_boot(NrState *nr, int n, cudaStream_t stream);
cudaError_t launch_twinpeaks_rearm(ChanState *state, int first, int count, cudaStream_t stream);
// the twin-peaks detector over the blocks of one call, before the receiver kernels (configure.cu)
cudaError_t launch_twinpeaks(const ChanParams *params, ChanState *state, const void *iq, int nch, int nblocks, long long chan_stride, cudaStream_t stream);

// spectrum ring of channels that ran on the tensor-core kernel (last 16 blocks of the call, factors from a.iqc_log)
cudaError_t launch_spectrum_tap(const RxArgs &a, cudaStream_t stream);
// the fused chain but for its deferred consumers: spectral NR hand-off and / or the (unzoomed) spectrum ring, both served around the
// tensor-core kernel
bool fused_eligible_ext(const ChanParams &p);
// UiSpectrum_RedrawSpectrum states 0-2
cudaError_t launch_spectrum(const ChanParams *params, const ChanState *state, const float *spec_ring, const float *pool,
                            int window_off, int twiddle_off, int first, int count, float *mags, cudaStream_t stream);

// UiSpectrum_RedrawSpectrum states 0-4 (spectrum.cu): display settings derived on the host (engine.cu)
struct SpecDisp {
    float filt_factor;     // 1 / ts.spectrum_filter, ui_spectrum.c:1434
    float db_scale;        // sd.db_scale, :1027-1036
    float agc_rate;        // sd.agc_rate, :988
    float cons;            // ts.dbm_constant - 225 - 3, :2004
    int scope_w;           // slayout.scope.w
};
cudaError_t launch_spectrum_display(const ChanParams *params, const ChanState *state, const float *spec_ring, const float *pool,
                                    int window_off, int twiddle_off, int first, int count, const SpecDisp &dc, float *avg_state,
                                    float *off_state, float *mags_out, float *avg_out, float *disp_out, float *lvl_out, cudaStream_t stream);

// configure: write params for channels [first, first+count) and apply the reference's state
// reset rules (reset != 0: boot state; 0: AudioDriver_SetProcessingChain semantics)
cudaError_t launch_configure(ChanParams *params, ChanState *state, NrState *nr, float *spec_ring, TxState *tx,
                             TxParams *txparams, const ChanParams &newp, const TxParams &newtx, int first, int count,
                             int stride, int reset, cudaStream_t stream);

}  // namespace uhsdr
