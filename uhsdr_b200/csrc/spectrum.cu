// spectrum.cu -- spectrum-display FFT (UiSpectrum_RedrawSpectrum states 0-2, ui_spectrum.c:1362-1390).
#include "dsp_device.cuh"
#include "kernels.h"

namespace uhsdr {
cudaError_t launch_spectrum(const ChanParams *params, const ChanState *state, const float *spec_ring, const float *pool,
                            int window_off, int twiddle_off, int first, int count, float *mags, cudaStream_t stream)
{
    (void)params; (void)state; (void)spec_ring; (void)pool; (void)window_off; (void)twiddle_off; (void)first; (void)count; (void)mags; (void)stream;
    return cudaErrorNotSupported;
}
}  // namespace uhsdr
