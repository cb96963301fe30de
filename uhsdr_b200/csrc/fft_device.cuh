// fft_device.cuh -- warp-cooperative in-place complex FFT in shared memory (decimation in time, two radix-2
// stages per pass).  Stands in for CMSIS arm_cfft_f32 (TransformFunctions/arm_cfft_f32.c:574-630): forward
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
    // Two radix-2 stages per pass (radix-2^2, in registers): half the passes, barriers and twiddle loads of a plain radix-2 walk.
    // Stage s (span h = 2^(s-1)) pairs (a, b) and (c, d) with W^(k N / 2h); stage s + 1 pairs (a', c') with w2 = W^(k N / 4h) and
    // (b', d') with W^((k + h) N / 4h) = -i w2.
    int s = 1;
    if (LOG2N & 1) {
        // odd number of stages: the first one alone (twiddle 1)
        for (int b = tid; b < N / 2; b += NTHR) {
            const int i = 2 * b, j = i + 1;
            const float ur = buf[2 * i], ui = buf[2 * i + 1], xr = buf[2 * j], xi = buf[2 * j + 1];
            buf[2 * j] = __fsub_rn(ur, xr); buf[2 * j + 1] = __fsub_rn(ui, xi);
            buf[2 * i] = __fadd_rn(ur, xr); buf[2 * i + 1] = __fadd_rn(ui, xi);
        }
        if (NTHR > 32) __syncthreads(); else __syncwarp();
        s = 2;
    }
#pragma unroll 1
    for (; s < LOG2N; s += 2) {
        const int h = 1 << (s - 1);
        for (int q = tid; q < N / 4; q += NTHR) {
            const int k = q & (h - 1);
            const int i0 = ((q >> (s - 1)) << (s + 1)) + k;
            const float w1r = __ldg(tw + 2 * (k << (LOG2N - s))), w1i = __ldg(tw + 2 * (k << (LOG2N - s)) + 1);
            const float w2r = __ldg(tw + 2 * (k << (LOG2N - s - 1))), w2i = __ldg(tw + 2 * (k << (LOG2N - s - 1)) + 1);
            float2 va = *reinterpret_cast<float2 *>(buf + 2 * i0), vb = *reinterpret_cast<float2 *>(buf + 2 * (i0 + h));
            float2 vc = *reinterpret_cast<float2 *>(buf + 2 * (i0 + 2 * h)), vd = *reinterpret_cast<float2 *>(buf + 2 * (i0 + 3 * h));
            // stage s
            float tr = __fsub_rn(__fmul_rn(vb.x, w1r), __fmul_rn(vb.y, w1i)), ti = __fadd_rn(__fmul_rn(vb.x, w1i), __fmul_rn(vb.y, w1r));
            const float ar = __fadd_rn(va.x, tr), ai = __fadd_rn(va.y, ti), br = __fsub_rn(va.x, tr), bi = __fsub_rn(va.y, ti);
            tr = __fsub_rn(__fmul_rn(vd.x, w1r), __fmul_rn(vd.y, w1i)); ti = __fadd_rn(__fmul_rn(vd.x, w1i), __fmul_rn(vd.y, w1r));
            const float cr = __fadd_rn(vc.x, tr), ci = __fadd_rn(vc.y, ti), dr = __fsub_rn(vc.x, tr), di = __fsub_rn(vc.y, ti);
            // stage s + 1
            tr = __fsub_rn(__fmul_rn(cr, w2r), __fmul_rn(ci, w2i)); ti = __fadd_rn(__fmul_rn(cr, w2i), __fmul_rn(ci, w2r));
            va.x = __fadd_rn(ar, tr); va.y = __fadd_rn(ai, ti); vc.x = __fsub_rn(ar, tr); vc.y = __fsub_rn(ai, ti);
            // (-i w2) d' = (w2i dr' ... ): -i (x + i y) = y - i x
            const float er = __fsub_rn(__fmul_rn(dr, w2r), __fmul_rn(di, w2i)), ei = __fadd_rn(__fmul_rn(dr, w2i), __fmul_rn(di, w2r));
            tr = ei; ti = -er;
            vb.x = __fadd_rn(br, tr); vb.y = __fadd_rn(bi, ti); vd.x = __fsub_rn(br, tr); vd.y = __fsub_rn(bi, ti);
            *reinterpret_cast<float2 *>(buf + 2 * i0) = va; *reinterpret_cast<float2 *>(buf + 2 * (i0 + h)) = vb;
            *reinterpret_cast<float2 *>(buf + 2 * (i0 + 2 * h)) = vc; *reinterpret_cast<float2 *>(buf + 2 * (i0 + 3 * h)) = vd;
        }
        if (NTHR > 32) __syncthreads(); else __syncwarp();
    }
    if (inverse) {
        const float sc = 1.0f / (float)N;
        for (int i = tid; i < N; i += NTHR) { buf[2 * i] = __fmul_rn(buf[2 * i], sc); buf[2 * i + 1] = __fmul_rn(-buf[2 * i + 1], sc); }
        if (NTHR > 32) __syncthreads(); else __syncwarp();
    }
}

// One warp, data in registers: lane L holds the E = N / 32 consecutive elements E L .. E L + E - 1 of the bit-reversed sequence
// (decimation in time).  The first log2(E) stages pair elements of one lane; the five last stages pair lane L with lane L ^ 2^j on
// the same register, so they are shuffles -- no shared-memory pass, no bank conflicts, no barrier.  Same twiddle table and the same
// butterflies as fft_inplace.  xr / xi: in = element brev(E L + r) of the signal (see fft_warp_load), out = bin E L + r.
template <int N, int LOG2N>
__device__ __forceinline__ void fft_warp(float (&xr)[N / 32], float (&xi)[N / 32], const float *__restrict__ tw, int lane)
{
    constexpr int E = N / 32, LE = LOG2N - 5;
#pragma unroll
    for (int s = 1; s <= LE; s++) {
        const int h = 1 << (s - 1);
#pragma unroll
        for (int r = 0; r < E; r++) {
            if (r & h) continue;
            const int k = r & (h - 1);
            const float2 w_ = *reinterpret_cast<const float2 *>(tw + 2 * (k << (LOG2N - s)));
            const float wr = w_.x, wi = w_.y;
            const float tr = __fsub_rn(__fmul_rn(xr[r + h], wr), __fmul_rn(xi[r + h], wi));
            const float ti = __fadd_rn(__fmul_rn(xr[r + h], wi), __fmul_rn(xi[r + h], wr));
            xr[r + h] = __fsub_rn(xr[r], tr); xi[r + h] = __fsub_rn(xi[r], ti);
            xr[r] = __fadd_rn(xr[r], tr); xi[r] = __fadd_rn(xi[r], ti);
        }
    }
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int s = LE + 1 + j;
        const bool upper = (lane >> j) & 1;
        const int kl = E * (lane & ((1 << j) - 1));
#pragma unroll
        for (int r = 0; r < E; r++) {
            const int tidx = (kl + r) << (LOG2N - s);
            const float2 w_ = *reinterpret_cast<const float2 *>(tw + 2 * tidx);        // tw: global or shared memory
            const float wr = w_.x, wi = w_.y;
            const float pr = __shfl_xor_sync(0xffffffffu, xr[r], 1 << j), pi = __shfl_xor_sync(0xffffffffu, xi[r], 1 << j);
            const float br = upper ? xr[r] : pr, bi = upper ? xi[r] : pi;          // the element of the upper half of the pair
            const float ar = upper ? pr : xr[r], ai = upper ? pi : xi[r];
            const float tr = __fsub_rn(__fmul_rn(br, wr), __fmul_rn(bi, wi));
            const float ti = __fadd_rn(__fmul_rn(br, wi), __fmul_rn(bi, wr));
            xr[r] = upper ? __fsub_rn(ar, tr) : __fadd_rn(ar, tr);
            xi[r] = upper ? __fsub_rn(ai, ti) : __fadd_rn(ai, ti);
        }
    }
}

// index into the signal of register r of lane `lane` (bit reversal of E lane + r over LOG2N bits)
template <int N, int LOG2N> __device__ __forceinline__ int fft_warp_src(int lane, int r)
{
    return (int)(__brev((unsigned)((N / 32) * lane + r)) >> (32 - LOG2N));
}

// fft_inplace's contract (interleaved re, im in shared memory, one warp) on the register transform: one bit-reversed load and one
// store pass instead of a pass per stage pair.
template <int N, int LOG2N>
__device__ __forceinline__ void fft_warp_smem(float *buf /* [2N] re,im */, const float *__restrict__ tw, bool inverse, int lane)
{
    constexpr int E = N / 32;
    float xr[E], xi[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const float2 v = *reinterpret_cast<const float2 *>(buf + 2 * fft_warp_src<N, LOG2N>(lane, r));
        xr[r] = v.x; xi[r] = inverse ? -v.y : v.y;
    }
    __syncwarp();
    fft_warp<N, LOG2N>(xr, xi, tw, lane);
    const float sc = 1.0f / (float)N;
#pragma unroll
    for (int r = 0; r < E; r++) {
        float2 v = make_float2(xr[r], xi[r]);
        if (inverse) { v.x = __fmul_rn(v.x, sc); v.y = __fmul_rn(-v.y, sc); }
        *reinterpret_cast<float2 *>(buf + 2 * (E * lane + r)) = v;
    }
    __syncwarp();
}

}  // namespace uhsdr
