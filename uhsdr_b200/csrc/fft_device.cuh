// fft_device.cuh -- warp-cooperative in-place complex FFT in shared memory (radix-2, decimation in
// time).  Stands in for CMSIS arm_cfft_f32 (TransformFunctions/arm_cfft_f32.c:574-630): forward
// = plain DFT; inverse = conjugate, forward transform, conjugate, scale by 1/N (:616-627).
// Twiddles e^{-2 pi i k / N} (k < N/2, interleaved cos, sin) come from the coefficient pool, where
// the host computed them in double precision.  Results agree with the reference's radix-8 kernel
// to float rounding ("need only match to tolerance, not bit order").
#pragma once
#include "dsp_device.cuh"

namespace uhsdr {

template <int N, int LOG2N, int NTHR>
__device__ __forceinline__ void fft_inplace(float *buf /* [2N] re,im */, const float *__restrict__ tw, bool inverse, int tid)
{
    // bit-reversal permutation (+ conjugate on the way in for the inverse)
    for (int i = tid; i < N; i += NTHR) {
        const int j = (int)(__brev((unsigned)i) >> (32 - LOG2N));
        if (i < j) {
            const float ar = buf[2 * i], ai = buf[2 * i + 1], br = buf[2 * j], bi = buf[2 * j + 1];
            buf[2 * i] = br; buf[2 * i + 1] = inverse ? -bi : bi;
            buf[2 * j] = ar; buf[2 * j + 1] = inverse ? -ai : ai;
        } else if (i == j && inverse) {
            buf[2 * i + 1] = -buf[2 * i + 1];
        }
    }
    if (NTHR > 32) __syncthreads(); else __syncwarp();
#pragma unroll 1
    for (int s = 1; s <= LOG2N; s++) {
        const int half = 1 << (s - 1);
        for (int b = tid; b < N / 2; b += NTHR) {
            const int k = b & (half - 1);
            const int i = ((b >> (s - 1)) << s) + k;
            const int j = i + half;
            const int tidx = k << (LOG2N - s);            // twiddle index: k * N / len
            const float wr = __ldg(tw + 2 * tidx), wi = __ldg(tw + 2 * tidx + 1);
            const float xr = buf[2 * j], xi = buf[2 * j + 1];
            const float tr = __fsub_rn(__fmul_rn(xr, wr), __fmul_rn(xi, wi));
            const float ti = __fadd_rn(__fmul_rn(xr, wi), __fmul_rn(xi, wr));
            const float ur = buf[2 * i], ui = buf[2 * i + 1];
            buf[2 * j] = __fsub_rn(ur, tr); buf[2 * j + 1] = __fsub_rn(ui, ti);
            buf[2 * i] = __fadd_rn(ur, tr); buf[2 * i + 1] = __fadd_rn(ui, ti);
        }
        if (NTHR > 32) __syncthreads(); else __syncwarp();
    }
    if (inverse) {
        const float sc = 1.0f / (float)N;
        for (int i = tid; i < N; i += NTHR) { buf[2 * i] = __fmul_rn(buf[2 * i], sc); buf[2 * i + 1] = __fmul_rn(-buf[2 * i + 1], sc); }
        if (NTHR > 32) __syncthreads(); else __syncwarp();
    }
}

}  // namespace uhsdr
