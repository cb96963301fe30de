// nr_device.cuh -- spectral noise reduction (audio_nr.c), warp-cooperative. Filled in below.
#pragma once
#include "dsp_device.cuh"
#include "demod_device.cuh"

namespace uhsdr {
__device__ inline void nr_block(const ChanParams &p, NrState &nr, const float *__restrict__ pool, float *buf, int n, float *scr, int lane)
{
    (void)p; (void)nr; (void)pool; (void)buf; (void)n; (void)scr; (void)lane;
}
}  // namespace uhsdr
