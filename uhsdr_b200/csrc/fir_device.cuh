// fir_device.cuh -- register-blocked FP32 FIR building blocks of the split general path (north_star form (1): shared-memory
// staging of the sample windows, float4 accesses, register-blocked taps).
//
// The CMSIS kernels compute y[m] = sum_k c[k] x[mM - (N-1) + k], k ascending (arm_fir_f32.c:522-529,
// arm_fir_decimate_f32.c:455-486).  Here the taps of a filter are staged once per launch in shared memory, FRONT-PADDED with
// zeros so that the input window of every output starts on a 16-byte boundary of the [history | new samples] buffer (the
// history length H is a multiple of 4): with pf = (4 - (N-1) % 4) % 4 leading zeros the window of output m starts at
// x[H + mM - (N-1+pf)], and the padded tap count NP = N + pf rounded up to a multiple of 4 is walked in groups of 4 taps with
// one LDS.128 of taps (broadcast) and one LDS.128 of samples per group.  Zero taps contribute acc + 0 * x = acc exactly, and
// the real taps are still visited in ascending order, so the exact build (mad = multiply, then add) stays bit-identical to the
// reference; the shipping build fuses.
#pragma once
#include "dsp_device.cuh"

namespace uhsdr {

__host__ __device__ constexpr int fir_pad_front(int ntaps) { return (4 - ((ntaps - 1) & 3)) & 3; }
__host__ __device__ constexpr int fir_padded_len(int ntaps) { return (ntaps + fir_pad_front(ntaps) + 3) & ~3; }

// taps -> shared memory in the padded layout (all lanes of the warp)
__device__ __forceinline__ void fir_stage_taps(float *dst, const float *__restrict__ c, int ntaps, int lane)
{
    const int pf = fir_pad_front(ntaps), np = fir_padded_len(ntaps);
    for (int i0 = 0; i0 < np; i0 += 128) {               // four loads in flight per lane
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = i0 + 32 * u + lane - pf;
            v[u] = (k >= 0 && k < ntaps) ? __ldg(c + k) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + 32 * u + lane;
            if (i < np) dst[i] = v[u];
        }
    }
}

template <bool EXACTSUM> __device__ __forceinline__ float fir_mac(float x, float c, float acc)
{
    if constexpr (EXACTSUM) return __fadd_rn(acc, __fmul_rn(x, c));
    else return mad(x, c, acc);
}

// M = 1, four consecutive outputs per lane: y[r] = sum_k t[k] x[start + r + k], r = 0..3; start % 4 == 0.
// One group of 4 taps: 1 LDS.128 of samples + 1 LDS.128 of taps + 16 multiply-adds.
template <bool EXACTSUM>
__device__ __forceinline__ void fir4_m1(const float *__restrict__ x, const float *__restrict__ t, int groups, float (&y)[4])
{
    float4 w = *reinterpret_cast<const float4 *>(x);
#pragma unroll
    for (int r = 0; r < 4; r++) y[r] = 0.0f;
#pragma unroll 2
    for (int g = 0; g < groups; g++) {
        const float4 n = *reinterpret_cast<const float4 *>(x + 4 * g + 4);
        const float4 c = *reinterpret_cast<const float4 *>(t + 4 * g);
        const float wv[8] = { w.x, w.y, w.z, w.w, n.x, n.y, n.z, n.w };
        const float cv[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) y[r] = fir_mac<EXACTSUM>(wv[r + j], cv[j], y[r]);
        w = n;
    }
}

// Two independent filters side by side (I with its taps, Q with its taps): eight accumulator chains and four loads in flight
// per group instead of four and two -- these loops wait on shared-memory latency at the occupancy the staging buffers allow.
template <bool EXACTSUM>
__device__ __forceinline__ void fir4_m1_pair(const float *__restrict__ xa, const float *__restrict__ ta, const float *__restrict__ xb,
                                             const float *__restrict__ tb, int groups, float (&ya)[4], float (&yb)[4])
{
    float4 wa = *reinterpret_cast<const float4 *>(xa), wb = *reinterpret_cast<const float4 *>(xb);
#pragma unroll
    for (int r = 0; r < 4; r++) { ya[r] = 0.0f; yb[r] = 0.0f; }
#pragma unroll 4
    for (int g = 0; g < groups; g++) {
        const float4 na = *reinterpret_cast<const float4 *>(xa + 4 * g + 4), nb = *reinterpret_cast<const float4 *>(xb + 4 * g + 4);
        const float4 ca = *reinterpret_cast<const float4 *>(ta + 4 * g), cb = *reinterpret_cast<const float4 *>(tb + 4 * g);
        const float va[8] = { wa.x, wa.y, wa.z, wa.w, na.x, na.y, na.z, na.w }, vb[8] = { wb.x, wb.y, wb.z, wb.w, nb.x, nb.y, nb.z, nb.w };
        const float ka[4] = { ca.x, ca.y, ca.z, ca.w }, kb[4] = { cb.x, cb.y, cb.z, cb.w };
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) { ya[r] = fir_mac<EXACTSUM>(va[r + j], ka[j], ya[r]); yb[r] = fir_mac<EXACTSUM>(vb[r + j], kb[j], yb[r]); }
        wa = na; wb = nb;
    }
}

// Same with two tap sets on one input (TX Hilbert pair: one microphone signal, I and Q filters).
template <bool EXACTSUM>
__device__ __forceinline__ void fir4_m1_dual(const float *__restrict__ x, const float *__restrict__ ta, const float *__restrict__ tb, int groups,
                                             float (&ya)[4], float (&yb)[4])
{
    float4 w = *reinterpret_cast<const float4 *>(x);
#pragma unroll
    for (int r = 0; r < 4; r++) { ya[r] = 0.0f; yb[r] = 0.0f; }
#pragma unroll 2
    for (int g = 0; g < groups; g++) {
        const float4 n = *reinterpret_cast<const float4 *>(x + 4 * g + 4);
        const float4 ca = *reinterpret_cast<const float4 *>(ta + 4 * g), cb = *reinterpret_cast<const float4 *>(tb + 4 * g);
        const float wv[8] = { w.x, w.y, w.z, w.w, n.x, n.y, n.z, n.w };
        const float av[4] = { ca.x, ca.y, ca.z, ca.w }, bv[4] = { cb.x, cb.y, cb.z, cb.w };
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) { ya[r] = fir_mac<EXACTSUM>(wv[r + j], av[j], ya[r]); yb[r] = fir_mac<EXACTSUM>(wv[r + j], bv[j], yb[r]); }
        w = n;
    }
}

// Decimation by 4, R outputs per lane 32 apart (m = lane + 32 r): y[r] = sum_k t[k] x[start + 128 r + k], start % 4 == 0
// (start = window of output m = lane).  One group: R LDS.128 of samples + 1 LDS.128 of taps + 4 R multiply-adds.
template <bool EXACTSUM, int R>
__device__ __forceinline__ void fir_dec4(const float *__restrict__ x, const float *__restrict__ t, int groups, float (&y)[R])
{
#pragma unroll
    for (int r = 0; r < R; r++) y[r] = 0.0f;
#pragma unroll 2
    for (int g = 0; g < groups; g++) {
        const float4 c = *reinterpret_cast<const float4 *>(t + 4 * g);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float4 v = *reinterpret_cast<const float4 *>(x + 128 * r + 4 * g);
            y[r] = fir_mac<EXACTSUM>(v.x, c.x, y[r]); y[r] = fir_mac<EXACTSUM>(v.y, c.y, y[r]);
            y[r] = fir_mac<EXACTSUM>(v.z, c.z, y[r]); y[r] = fir_mac<EXACTSUM>(v.w, c.w, y[r]);
        }
    }
}

// General strided form (any M, scalar sample loads, taps still staged and padded): y = sum_k t[k] x[k]; used where the window
// start is not 16-byte aligned (decimation by 2).
template <bool EXACTSUM>
__device__ __forceinline__ float fir_scalar(const float *__restrict__ x, const float *__restrict__ t, int np)
{
    float acc = 0.0f;
    for (int k = 0; k < np; k += 4) {
        const float4 c = *reinterpret_cast<const float4 *>(t + k);
        acc = fir_mac<EXACTSUM>(x[k], c.x, acc); acc = fir_mac<EXACTSUM>(x[k + 1], c.y, acc);
        acc = fir_mac<EXACTSUM>(x[k + 2], c.z, acc); acc = fir_mac<EXACTSUM>(x[k + 3], c.w, acc);
    }
    return acc;
}

}  // namespace uhsdr
