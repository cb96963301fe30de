// spectrum.cu -- spectrum-display FFT: UiSpectrum_RedrawSpectrum states 0-2
// (mchf-eclipse/drivers/ui/lcd/ui_spectrum.c:1362-1390) for the 480x320 layout (fft_iq_len 1024,
// 512-point complex FFT, :975-979).  One CTA of 128 threads per channel:
//   state 0  snapshot of the spectrum ring starting at samp_ptr (ring order Q,I,Q,I ...), multiply
//            EVERY float by von_Hann_1024[i] (:409-413), scale by 1/ads.codec_gain_calc (:438-439)
//   state 1  arm_cfft_f32 length 512
//   state 2  arm_cmplx_mag_f32 -> 512 magnitudes
#include "fft_device.cuh"
#include "kernels.h"

namespace uhsdr {

__global__ void __launch_bounds__(128)
spectrum_kernel(const ChanParams *__restrict__ params, const ChanState *__restrict__ state, const float *__restrict__ spec_ring,
                const float *__restrict__ pool, int window_off, int twiddle_off, int first, int count, float *__restrict__ mags)
{
    // one warp per channel, the whole transform in registers (fft_warp: 16 elements per lane, the last five stages by shuffle);
    // the twiddle table (256 complex) staged once per CTA in shared memory
    __shared__ __align__(16) float tws[512];
    for (int i = threadIdx.x; i < 512; i += 128) tws[i] = __ldg(pool + twiddle_off + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slot >= count) return;
    const int ch = first + slot;
    const ChanParams &p = params[ch];
    float *out = mags + (size_t)slot * 512;
    if (!p.configured || !p.spectrum_enable) {
        for (int i = lane; i < 512; i += 32) out[i] = 0.0f;
        return;
    }
    const uint32_t sp = state[ch].samp_ptr;
    const float *ring = spec_ring + (size_t)ch * 1024;
    const float *win = pool + window_off;
    const float gcalc = (float)(1.0 / (double)p.codec_gain_calc);
    float xr[16], xi[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        // element n of the 512 complex samples = floats 2n, 2n + 1 of the snapshot; the window runs over the interleaved array
        const int n = fft_warp_src<512, 9>(lane, r);
        xr[r] = __fmul_rn(__fmul_rn(ring[(sp + 2u * (uint32_t)n) & 1023u], __ldg(win + 2 * n)), gcalc);
        xi[r] = __fmul_rn(__fmul_rn(ring[(sp + 2u * (uint32_t)n + 1u) & 1023u], __ldg(win + 2 * n + 1)), gcalc);
    }
    fft_warp<512, 9>(xr, xi, tws, lane);
    float4 *o4 = reinterpret_cast<float4 *>(out + 16 * lane);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        float m[4];
#pragma unroll
        for (int e = 0; e < 4; e++) m[e] = __fsqrt_rn(__fadd_rn(__fmul_rn(xr[4 * q + e], xr[4 * q + e]), __fmul_rn(xi[4 * q + e], xi[4 * q + e])));
        if (((uintptr_t)out & 15) == 0) o4[q] = make_float4(m[0], m[1], m[2], m[3]);
        else { for (int e = 0; e < 4; e++) out[16 * lane + 4 * q + e] = m[e]; }
    }
}

// UiSpectrum_RedrawSpectrum states 0-4 for one channel per CTA: the FFT above, then
//   state 3  IIR bin averaging (ui_spectrum.c:1432-1446: avg -= avg/f; avg += mag/f; floor 1) and UiSpectrum_CalculateDBm
//            (:1990-2122: sum of the frequency-ordered magnitudes x SCOPE_PREAMP_GAIN over the passband bins,
//            19.8 log10f_fast(sum) + cons; the dBm/Hz figure subtracts 10 log10f_fast of the passband width)
//   state 4  UiSpectrum_ScaleFFT (:1258-1296: display_offset + log10f_fast(avg) * db_scale in frequency order, floor 1, running
//            minimum), UiSpectrum_ScaleFFT2SpectrumWidth (:1300-1337) when the scope is narrower than 512 bins, and the
//            sliding display offset (:1485).
// avg_state / off_state are the channel's FFT_AVGData and sd.display_offset.  Sums that the reference forms sequentially
// (passband sum, width rescaling) are formed sequentially by one thread: they are a few hundred additions per call.
__global__ void __launch_bounds__(128)
spectrum_display_kernel(const ChanParams *__restrict__ params, const ChanState *__restrict__ state, const float *__restrict__ spec_ring,
                        const float *__restrict__ pool, int window_off, int twiddle_off, int first, SpecDisp dc,
                        float *__restrict__ avg_state, float *__restrict__ off_state,
                        float *__restrict__ mags_out, float *__restrict__ avg_out, float *__restrict__ disp_out, float *__restrict__ lvl_out)
{
    __shared__ __align__(16) float buf[1024];
    __shared__ float mag[512];
    __shared__ float red[4];
    const int ch = first + blockIdx.x;
    const int tid = threadIdx.x;
    const ChanParams &p = params[ch];
    float *disp = disp_out + (size_t)blockIdx.x * dc.scope_w;
    float *lvl = lvl_out + (size_t)blockIdx.x * 3;
    if (!p.configured || !p.spectrum_enable) {
        for (int i = tid; i < dc.scope_w; i += 128) disp[i] = 0.0f;
        if (mags_out) for (int i = tid; i < 512; i += 128) mags_out[(size_t)blockIdx.x * 512 + i] = 0.0f;
        if (avg_out) for (int i = tid; i < 512; i += 128) avg_out[(size_t)blockIdx.x * 512 + i] = 0.0f;
        if (tid < 3) lvl[tid] = 0.0f;
        return;
    }
    // ---- states 0-2 ----
    const uint32_t sp = state[ch].samp_ptr;
    const float *ring = spec_ring + (size_t)ch * 1024;
    const float gcalc = (float)(1.0 / (double)p.codec_gain_calc);
    for (int i = tid; i < 1024; i += 128) {
        const float v = ring[(sp + (uint32_t)i) & 1023u];
        buf[i] = __fmul_rn(__fmul_rn(v, __ldg(pool + window_off + i)), gcalc);
    }
    __syncthreads();
    fft_inplace<512, 9, 128>(buf, pool + twiddle_off, false, tid);
    for (int i = tid; i < 512; i += 128) {
        const float re = buf[2 * i], im = buf[2 * i + 1];
        const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
        mag[i] = m;
        if (mags_out) mags_out[(size_t)blockIdx.x * 512 + i] = m;
    }
    __syncthreads();
    // ---- state 3: averaging ----
    float *avg = avg_state + (size_t)ch * 512;
    float *fs = buf;                       // sd.FFT_Samples
    float av[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = tid + 128 * k;
        float a = avg[i];
        a = __fsub_rn(a, __fmul_rn(a, dc.filt_factor));
        a = __fadd_rn(__fmul_rn(mag[i], dc.filt_factor), a);
        if (a < 1.0f) a = 1.0f;
        avg[i] = a;
        if (avg_out) avg_out[(size_t)blockIdx.x * 512 + i] = a;
        av[k] = a;
        // CalculateDBm's frequency-ordered copy of the NEW magnitudes (:2084-2091)
        const int src = i < 256 ? i + 256 : i - 256;
        fs[512 - i - 1] = __fmul_rn(mag[src], 1000.0f);                 // SCOPE_PREAMP_GAIN
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) mag[tid + 128 * k] = av[k];             // averaged data for state 4
    __syncthreads();
    if (tid == 0) {
        float sum_db = 0.0f;
        for (int c = p.dbm_lbin; c <= p.dbm_ubin; c++) sum_db = __fadd_rn(sum_db, fs[c]);
        float dbm = -145.0f, dbmhz = -145.0f;
        if (sum_db > 0.0f) {
            dbm = __fadd_rn(__fmul_rn(19.8f, log10f_fast(sum_db)), dc.cons);
            dbmhz = __fsub_rn(dbm, __fmul_rn(10.0f, log10f_fast(p.dbm_span_hz)));
        }
        lvl[0] = dbm; lvl[1] = dbmhz;
    }
    __syncthreads();
    // ---- state 4: log scaling in frequency order ----
    const float doff = off_state[ch];
    float mn = 100000.0f;
    for (int i = tid; i < 512; i += 128) {
        const int src = i < 256 ? i + 256 : i - 256;
        const float sig = __fadd_rn(doff, __fmul_rn(log10f_fast(mag[src]), dc.db_scale));
        mn = fminf(mn, sig);
        fs[512 - i - 1] = sig < 1.0f ? 1.0f : sig;
    }
    for (int d = 16; d > 0; d >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    if ((tid & 31) == 0) red[tid >> 5] = mn;
    __syncthreads();
    if (tid == 0) {
        const float min1 = fminf(fminf(red[0], red[1]), fminf(red[2], red[3]));
        if (dc.scope_w != 512) {
            // UiSpectrum_ScaleFFT2SpectrumWidth, in place, in the reference's order
            const float full_amount = __fdiv_rn(512.0f, (float)dc.scope_w);
            float amount = full_amount, value = 0.0f;
            int idx_new = 0, idx_old = 0;
            do {
                while (amount >= 1.0f) { value = __fadd_rn(value, fs[idx_old]); idx_old++; amount = (float)((double)amount - 1.0); }
                const float for_next = __fmul_rn(__fsub_rn(1.0f, amount), fs[idx_old]);
                value = __fadd_rn(value, __fsub_rn(fs[idx_old], for_next));
                fs[idx_new] = value;
                idx_new++;
                if (idx_new < dc.scope_w) { value = for_next; idx_old++; amount = __fsub_rn(full_amount, __fsub_rn(1.0f, amount)); }
            } while (idx_new < dc.scope_w);
        }
        const float noff = __fsub_rn(doff, __fdiv_rn(__fmul_rn(dc.agc_rate, min1), 5.0f));
        off_state[ch] = noff;
        lvl[2] = noff;
    }
    __syncthreads();
    for (int i = tid; i < dc.scope_w; i += 128) disp[i] = fs[i];
}

cudaError_t launch_spectrum_display(const ChanParams *params, const ChanState *state, const float *spec_ring, const float *pool,
                                    int window_off, int twiddle_off, int first, int count, const SpecDisp &dc, float *avg_state,
                                    float *off_state, float *mags_out, float *avg_out, float *disp_out, float *lvl_out, cudaStream_t stream)
{
    if (window_off < 0 || twiddle_off < 0 || dc.scope_w < 1 || dc.scope_w > 512) return cudaErrorInvalidValue;
    spectrum_display_kernel<<<count, 128, 0, stream>>>(params, state, spec_ring, pool, window_off, twiddle_off, first, dc, avg_state, off_state,
                                                       mags_out, avg_out, disp_out, lvl_out);
    return cudaGetLastError();
}

cudaError_t launch_spectrum(const ChanParams *params, const ChanState *state, const float *spec_ring, const float *pool,
                            int window_off, int twiddle_off, int first, int count, float *mags, cudaStream_t stream)
{
    if (window_off < 0 || twiddle_off < 0) return cudaErrorInvalidValue;
    spectrum_kernel<<<(count + 3) / 4, 128, 0, stream>>>(params, state, spec_ring, pool, window_off, twiddle_off, first, count, mags);
    return cudaGetLastError();
}

// Spectrum sample tap for channels that ran on the tensor-core kernel (whose front end keeps no copy of the corrected samples):
// AudioDriver_SpectrumNoZoomProcessSamples (audio_driver.c:1811-1849) writes every block's corrected, not yet translated (Q, I) pairs
// into the 1024-float ring, so after a call only its last 16 blocks are in it.  One warp per channel re-applies the IQ correction to
// those blocks (factors logged by the tensor-core kernel's front end, a.iqc_log) and advances samp_ptr by the whole call.
__global__ void __launch_bounds__(128)
spectrum_tap_kernel(RxArgs a)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slot >= a.num_items) return;
    const int ch = a.chan_list ? a.chan_list[slot] : slot;
    const ChanParams &p = a.params[ch];
    if (!p.spectrum_enable || !a.spec_ring || p.zoom_m != 0) return;
    ChanState &st = a.state[ch];
    const int ntap = min(16, a.nblocks);
    const uint32_t sp0 = st.samp_ptr;
    uint32_t sp = (sp0 + 64u * (uint32_t)(a.nblocks - ntap)) & 1023u;
    const int2 *iq = reinterpret_cast<const int2 *>(a.iq) + (size_t)ch * (size_t)a.chan_stride + (size_t)(a.nblocks - ntap) * BLK;
    float *ring = a.spec_ring + (size_t)ch * 1024;
    const int iq_auto = p.iq_auto;
    const float adj_i = p.adj_i, adj_q = p.adj_q, phase_bal = p.phase_bal;
    for (int j = 0; j < ntap; j++) {
        const int2 s = iq[(size_t)j * BLK + lane];
        float fi = __fmul_rn((float)s.x, 0.0000152587890625f), fq = __fmul_rn((float)s.y, 0.0000152587890625f);
        if (iq_auto) {
            const float2 c = a.iqc_log[(size_t)slot * 16 + j];
            fq = __fadd_rn(fq, __fmul_rn(c.x, fi));
            fi = __fmul_rn(fi, c.y);
        } else {
            fi = __fmul_rn(fi, adj_i); fq = __fmul_rn(fq, adj_q);
            if (phase_bal < 0.0f) fq = __fadd_rn(fq, __fmul_rn(fi, phase_bal));
            else if (phase_bal > 0.0f) fi = __fadd_rn(fi, __fmul_rn(fq, phase_bal));
        }
        uint32_t ptr = sp + 2u * (uint32_t)lane;
        if (ptr >= 1024u) ptr -= 1024u;
        ring[ptr] = fq; ring[ptr + 1] = fi;
        sp = (sp + 64u) & 1023u;
    }
    if (lane == 0) st.samp_ptr = (sp0 + 64u * (uint32_t)a.nblocks) & 1023u;
}

cudaError_t launch_spectrum_tap(const RxArgs &a, cudaStream_t stream)
{
    if (a.num_items <= 0 || a.nblocks <= 0) return cudaSuccess;
    if (a.iqc_log == nullptr) return cudaErrorInvalidValue;
    spectrum_tap_kernel<<<(a.num_items + 3) / 4, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace uhsdr
