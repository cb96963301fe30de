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
                const float *__restrict__ pool, int window_off, int twiddle_off, int first, float *__restrict__ mags)
{
    __shared__ __align__(16) float buf[1024];
    const int ch = first + blockIdx.x;
    const int tid = threadIdx.x;
    const ChanParams &p = params[ch];
    float *out = mags + (size_t)blockIdx.x * 512;
    if (!p.configured || !p.spectrum_enable) {
        for (int i = tid; i < 512; i += 128) out[i] = 0.0f;
        return;
    }
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
        out[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
    }
}

cudaError_t launch_spectrum(const ChanParams *params, const ChanState *state, const float *spec_ring, const float *pool,
                            int window_off, int twiddle_off, int first, int count, float *mags, cudaStream_t stream)
{
    if (window_off < 0 || twiddle_off < 0) return cudaErrorInvalidValue;
    spectrum_kernel<<<count, 128, 0, stream>>>(params, state, spec_ring, pool, window_off, twiddle_off, first, mags);
    return cudaGetLastError();
}

}  // namespace uhsdr
