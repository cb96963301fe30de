// configure.cu -- device side of uhsdr_configure_channels: stores the new parameter record and
// applies the reference's state rules.
//   reset != 0 : fresh-process state (SURVEY.md 8a "state inventory": everything zero except
//                agc out_index = -1, FM squelched, NCO vector (0,1), M_c2 = 1, ALC = 1, NR Hk = 1 ...).
//   always     : what AudioDriver_SetProcessingChain does (audio_driver.c:1093-1251): lattice
//                states, FIR / decimator / interpolator histories and the IQ-correction averages
//                are cleared; the AGC ring is re-initialised only when the decimated rate changed
//                (audio_agc.c:138-142) and in_index is re-derived from out_index (:292-293);
//                biquad, SAM, FM, fade-leveler and NR state are kept.
#include "dsp_device.cuh"
#include "kernels.h"

namespace uhsdr {

__device__ void nr_boot_state(NrState &n)
{
    // NR_Init (audio_nr.c:78-103) + first_time==1 branch (:1920-1935) happens lazily in the kernel
    n.first_time = 1; n.init_counter = 0; n.was_here = 0; n.out_buffer = -1;
}

__global__ void nr_boot_kernel(NrState *nr, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nr_boot_state(nr[i]);
}

__global__ void tx_boot_kernel(TxState *tx, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tx[i].alc_val = 1.0f;      // tx_processor.c:137
}

__global__ void configure_kernel(ChanParams *params, ChanState *state, NrState *nr, float *spec_ring, TxState *tx,
                                 TxParams *txparams, const __grid_constant__ ChanParams newp,
                                 const __grid_constant__ TxParams newtx, int first, int count, int stride, int reset)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int ch = first + i * stride;
    params[ch] = newp;
    if (txparams) txparams[ch] = newtx;
    ChanState &s = state[ch];
    if (reset) {
        uint32_t *w = reinterpret_cast<uint32_t *>(&s);
        for (int k = 0; k < (int)(sizeof(ChanState) / 4); k++) w[k] = 0u;
        s.osc_vect_q = 1.0f;            // freq_shift.c:45-46
        s.fm_squelched = 1;             // audio_driver.c:475
        s.agc_out_index = -1;           // audio_agc.c:190
        s.agc_sample_rate = 0.0f;
        s.tw_state = 2;                 // ts.twinpeaks_tested = TWINPEAKS_WAIT, uhsdr_main.c:339
        if (nr) {
            uint32_t *wn = reinterpret_cast<uint32_t *>(&nr[ch]);
            for (int k = 0; k < (int)(sizeof(NrState) / 4); k++) wn[k] = 0u;
            nr_boot_state(nr[ch]);
        }
        if (spec_ring) for (int k = 0; k < 1024; k++) spec_ring[(size_t)ch * 1024 + k] = 0.0f;
        if (tx) {
            uint32_t *wt = reinterpret_cast<uint32_t *>(&tx[ch]);
            for (int k = 0; k < (int)(sizeof(TxState) / 4); k++) wt[k] = 0u;
            tx[ch].alc_val = 1.0f;
        }
    }
    // AudioDriver_SetProcessingChain
    for (int k = 0; k < MAX_LAT; k++) { s.pre_s[k] = 0.0f; s.aa_s[k] = 0.0f; }
    s.M_c1 = 0.0f; s.M_c2 = 1.0f; s.teta1_old = 0.0f; s.teta2_old = 0.0f; s.teta3_old = 0.0f;
    for (int k = 0; k < H1; k++) { s.s1_hist_i[k] = 0.0f; s.s1_hist_q[k] = 0.0f; }
    for (int k = 0; k < H2; k++) { s.s2_hist_i[k] = 0.0f; s.s2_hist_q[k] = 0.0f; }
    for (int k = 0; k < INTERP_HIST; k++) s.interp_hist[k] = 0.0f;
    for (int k = 0; k < 4; k++) { s.zoom_hist_i[k] = 0.0f; s.zoom_hist_q[k] = 0.0f; }     // arm_fir_decimate_init_f32, audio_driver.c:1073-1086
    // auto-notch init (audio_driver.c:1165-1187): LMS state, energy and the delay line are cleared; the coefficients and
    // the delay-line positions (function statics of AudioDriver_NotchFilter) survive a reconfiguration
    for (int k = 0; k < NOTCH_TAPS; k++) s.notch_x[k] = 0.0f;
    for (int k = 0; k < NOTCH_DELAY; k++) s.notch_delay[k] = 0.0f;
    s.notch_energy = 0.0f; s.notch_x0 = 0.0f; s.notch_head = 0;
    if (s.agc_sample_rate != newp.agc.sample_rate) {
        s.agc_sample_rate = newp.agc.sample_rate;
        for (int k = 0; k < AGC_RB; k++) s.agc_ring[k] = 0.0f;
        s.agc_out_index = -1;
        s.agc_ring_max = 0.0f; s.agc_volts = 0.0f; s.agc_save_volts = 0.0f;
        s.agc_fast_backaverage = 0.0f; s.agc_hang_backaverage = 0.0f;
        s.agc_hang_counter = 0; s.agc_decay_type = 0; s.agc_state = 0;
    }
    s.agc_in_index = (int)((uint32_t)(newp.agc.attack_buffsize + s.agc_out_index) % (uint32_t)AGC_RB);
    if (tx) {
        // TxProcessor_Set (tx_processor.c:72-119): lattice state and Hilbert histories cleared
        for (int k = 0; k < MAX_LAT; k++) tx[ch].lat_s[k] = 0.0f;
        for (int k = 0; k < H2; k++) tx[ch].hist[k] = 0.0f;
        tx[ch].fm_dds_sub_acc = 0u; tx[ch].fm_dds_burst_acc = 0u;       // softdds_setFreqDDS, smooth == false (softdds.c:38-45)
    }
}

cudaError_t launch_configure(ChanParams *params, ChanState *state, NrState *nr, float *spec_ring, TxState *tx,
                             TxParams *txparams, const ChanParams &newp, const TxParams &newtx, int first, int count,
                             int stride, int reset, cudaStream_t stream)
{
    const int threads = 64;
    configure_kernel<<<(count + threads - 1) / threads, threads, 0, stream>>>(params, state, nr, spec_ring, tx, txparams,
                                                                               newp, newtx, first, count, stride, reset);
    return cudaGetLastError();
}

__global__ void twinpeaks_rearm_kernel(ChanState *state, int first, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) state[first + i].tw_state = 2;       // ui_driver.c:7425; the detector's own statics keep their values
}

// AudioDriver_RxHandleTwinpeaks as its own small kernel, launched in front of the receiver kernels while a channel's detector can
// still be active (the 1050 blocks after a reset or a re-arm): one warp per channel walks through the call's blocks, forms the
// block statistics of the automatic IQ correction from the raw input (audio_driver.c:2274-2283, reference summation order, EMA
// in double), and feeds teta1 / teta3 of every block to the detector.  It reads the correction averages the receiver kernel will
// start from and writes nothing but the detector's own state -- so none of the hot kernels carries the detector.
__global__ void __launch_bounds__(128)
twinpeaks_kernel(const ChanParams *__restrict__ params, ChanState *state, const void *iq, int nch, int nblocks, long long chan_stride)
{
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ch >= nch) return;
    const ChanParams &p = params[ch];
    ChanState *st = state + ch;
    TwinPeaks tw = twinpeaks_load(st);
    if (!p.configured || !p.iq_auto || !twinpeaks_active(tw)) return;
    const int2 *__restrict__ src = reinterpret_cast<const int2 *>(iq) + (size_t)ch * (size_t)chan_stride;
    float te1 = st->teta1_old, te2 = st->teta2_old, te3 = st->teta3_old;
    for (int b = 0; b < nblocks && twinpeaks_active(tw); b++) {
        const int2 s = src[(size_t)b * BLK + lane];
        const float fi = __fmul_rn((float)s.x, 0.0000152587890625f), fq = __fmul_rn((float)s.y, 0.0000152587890625f);
        const float t1 = __fmul_rn(sign_new(fi), fq), t2 = __fmul_rn(sign_new(fi), fi), t3 = __fmul_rn(sign_new(fq), fq);
        float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#if UHSDR_EXACT
        for (int j = 0; j < 32; j++) {                       // the reference's summation order
            s1 = __fadd_rn(s1, __shfl_sync(0xffffffffu, t1, j));
            s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, t2, j));
            s3 = __fadd_rn(s3, __shfl_sync(0xffffffffu, t3, j));
        }
#else
        s1 = t1; s2 = t2; s3 = t3;                           // butterfly: the detector's 22.5 degree threshold does not see the rounding
        for (int d = 16; d > 0; d >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, d);
            s2 += __shfl_xor_sync(0xffffffffu, s2, d);
            s3 += __shfl_xor_sync(0xffffffffu, s3, d);
        }
#endif
        te1 = (float)(-0.003 * (double)__fdiv_rn(s1, 32.0f) + 0.997 * (double)te1);
        te2 = (float)(0.003 * (double)__fdiv_rn(s2, 32.0f) + 0.997 * (double)te2);
        te3 = (float)(0.003 * (double)__fdiv_rn(s3, 32.0f) + 0.997 * (double)te3);
        tw = twinpeaks_block(tw, te1, te3);
    }
    if (lane == 0) twinpeaks_store(st, tw);
}

cudaError_t launch_twinpeaks(const ChanParams *params, ChanState *state, const void *iq, int nch, int nblocks, long long chan_stride, cudaStream_t stream)
{
    twinpeaks_kernel<<<(nch + 3) / 4, 128, 0, stream>>>(params, state, iq, nch, nblocks, chan_stride);
    return cudaGetLastError();
}

cudaError_t launch_twinpeaks_rearm(ChanState *state, int first, int count, cudaStream_t stream)
{
    twinpeaks_rearm_kernel<<<(count + 127) / 128, 128, 0, stream>>>(state, first, count);
    return cudaGetLastError();
}

cudaError_t launch_nr_boot(NrState *nr, int n, cudaStream_t stream)
{
    nr_boot_kernel<<<(n + 127) / 128, 128, 0, stream>>>(nr, n);
    return cudaGetLastError();
}

cudaError_t launch_tx_boot(TxState *tx, int n, cudaStream_t stream)
{
    tx_boot_kernel<<<(n + 127) / 128, 128, 0, stream>>>(tx, n);
    return cudaGetLastError();
}

}  // namespace uhsdr
