// dsp_device.cuh -- per-sample device building blocks of the receiver chain.
//
// Two arithmetic flavours, selected at compile time (the library is built twice):
//   UHSDR_EXACT=1  every sum in the reference's order with separate multiply and add
//                  (nvcc -fmad=false), so that integer formatting, decimation phase and all
//                  control decisions can be proven bit-exact against the oracle;
//   UHSDR_EXACT=0  (shipping build) fused multiply-adds and re-associated FIR sums; results are
//                  within the tolerance north_star states (|err| <= 1e-4 relative, >= 90 dB SNR).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "uhsdr_dev.h"

#ifndef UHSDR_EXACT
#define UHSDR_EXACT 0
#endif

namespace uhsdr {

// a*b + c: fused in the fast build, two roundings (reference arithmetic) in the exact build.
__device__ __forceinline__ float mad(float a, float b, float c)
{
#if UHSDR_EXACT
    return __fadd_rn(__fmul_rn(a, b), c);
#else
    return fmaf(a, b, c);
#endif
}

// Math_sign_new, misc/uhsdr_math.c:65-67
__device__ __forceinline__ float sign_new(float x) { return (x < 0.0f) ? -1.0f : ((x > 0.0f) ? 1.0f : 0.0f); }

// Math_log10f_fast, misc/uhsdr_math.c:27-41: cubic in the frexpf mantissa.
__device__ __forceinline__ float log10f_fast(float X)
{
    int E;
    float F = frexpf(fabsf(X), &E);
    float Y = 1.23149591368684f;
    Y = __fmul_rn(Y, F); Y = __fadd_rn(Y, -4.11852516267426f);
    Y = __fmul_rn(Y, F); Y = __fadd_rn(Y, 6.02197014179219f);
    Y = __fmul_rn(Y, F); Y = __fadd_rn(Y, -3.13396450166353f);
    Y = __fadd_rn(Y, (float)E);
    return __fmul_rn(Y, 0.3010299956639812f);
}

// arm_iir_lattice_f32 (CMSIS FilteringFunctions/arm_iir_lattice_f32.c:348-440), one sample.
// s[0..N-1]: g-state; k[N], v[N+1] coefficient arrays in stored order.
__device__ __forceinline__ float lattice_step(float x, float *s, const float *__restrict__ k,
                                              const float *__restrict__ v, int N)
{
    float f = x, acc = 0.0f, fn = 0.0f;
    for (int j = 0; j < N; j++) {
        float g = s[j];
        float kj = __ldg(k + j);
        fn = __fsub_rn(f, __fmul_rn(kj, g));
        float gn = __fadd_rn(__fmul_rn(fn, kj), g);
        acc = __fadd_rn(acc, __fmul_rn(gn, __ldg(v + j)));
        if (j > 0) s[j - 1] = gn;      // next state: s[j-1] = w[j]
        f = fn;
    }
    acc = __fadd_rn(acc, __fmul_rn(fn, __ldg(v + N)));
    if (N > 0) s[N - 1] = fn;
    return acc;
}

// arm_biquad_cascade_df1_f32 (arm_biquad_cascade_df1_f32.c:349-418), one stage, one sample.
__device__ __forceinline__ float biquad_step(float x, const float *c, BiquadS &st)
{
    float acc = __fmul_rn(c[0], x);
    acc = __fadd_rn(acc, __fmul_rn(c[1], st.x1));
    acc = __fadd_rn(acc, __fmul_rn(c[2], st.x2));
    acc = __fadd_rn(acc, __fmul_rn(c[3], st.y1));
    acc = __fadd_rn(acc, __fmul_rn(c[4], st.y2));
    st.x2 = st.x1; st.x1 = x; st.y2 = st.y1; st.y1 = acc;
    return acc;
}

// AudioAgc_RunAgcWdsp (audio_agc.c:367-575), one sample, mono.  The reference keeps a separate
// abs_ring; |ring[k]| is the same value, so only the sample ring is stored.
struct AgcRun {
    int out_index, in_index;
    float ring_max, volts, save_volts, fast_backaverage, hang_backaverage;
    int hang_counter, decay_type, state, action, hang_action;
};

__device__ __forceinline__ float agc_step(float x, const AgcP &a, AgcRun &r, float *ring)
{
    if (++r.out_index >= AGC_RB) r.out_index -= AGC_RB;
    if (++r.in_index >= AGC_RB) r.in_index -= AGC_RB;
    const float out_sample = ring[r.out_index];
    const float abs_out = fabsf(out_sample);
    ring[r.in_index] = x;
    const float abs_in = fabsf(x);

    r.fast_backaverage = __fadd_rn(__fmul_rn(a.fast_backmult, abs_out), __fmul_rn(a.onemfast_backmult, r.fast_backaverage));
    r.hang_backaverage = __fadd_rn(__fmul_rn(a.hang_backmult, abs_out), __fmul_rn(a.onemhang_backmult, r.hang_backaverage));
    r.hang_action = (r.hang_backaverage > a.hang_level) ? 1 : 0;

    if ((abs_out >= r.ring_max) && (abs_out > 0.0f)) {
        float m = 0.0f;
        int k = r.out_index;
        for (int j = 0; j < a.attack_buffsize; j++) {
            if (++k == AGC_RB) k = 0;
            m = fmaxf(m, fabsf(ring[k]));
        }
        r.ring_max = m;
    }
    if (abs_in > r.ring_max) r.ring_max = abs_in;
    if (r.hang_counter > 0) --r.hang_counter;

    const float d = __fsub_rn(r.ring_max, r.volts);
    const bool attack = r.ring_max >= r.volts;
    switch (r.state) {
    case 0:
        if (attack) {
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.attack_mult));
        } else if (r.volts > __fmul_rn(a.pop_ratio, r.fast_backaverage)) {
            r.state = 1;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.fast_decay_mult));
        } else if (a.hang_enable && (r.hang_backaverage > a.hang_level)) {
            r.state = 2;
            r.hang_counter = (int)__fmul_rn(a.hangtime, a.sample_rate);
            r.decay_type = 1;
        } else {
            r.state = 3;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.decay_mult));
            r.decay_type = 0;
        }
        break;
    case 1:
        if (attack) {
            r.state = 0;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.attack_mult));
        } else if (r.volts > r.save_volts) {
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.fast_decay_mult));
        } else if (r.hang_counter > 0) {
            r.state = 2;
        } else if (r.decay_type == 0) {
            r.state = 3;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.decay_mult));
        } else {
            r.state = 4;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.hang_decay_mult));
        }
        break;
    case 2:
        if (attack) {
            r.state = 0;
            r.save_volts = r.volts;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.attack_mult));
        } else if (r.hang_counter == 0) {
            r.state = 4;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.hang_decay_mult));
        }
        break;
    case 3:
        if (attack) {
            r.state = 0;
            r.save_volts = r.volts;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.attack_mult));
        } else {
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.decay_mult));
        }
        break;
    default:
        if (attack) {
            r.state = 0;
            r.save_volts = r.volts;
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.attack_mult));
        } else {
            r.volts = __fadd_rn(r.volts, __fmul_rn(d, a.hang_decay_mult));
        }
        break;
    }
    if (r.volts < a.min_volts) { r.volts = a.min_volts; r.action = 0; } else { r.action = 1; }

    float vo = log10f_fast(__fmul_rn(a.inv_max_input, r.volts));
    if (vo > 0.0f) vo = 0.0f;
    const float mult = __fdiv_rn(__fsub_rn(a.out_target, __fmul_rn(a.slope_constant, vo)), r.volts);
    return __fmul_rn(out_sample, mult);
}

// AudioDriver_RxHandleTwinpeaks (audio_driver.c:2173-2248), one block: called after the block's IQ-correction averages teta1 /
// teta3 have been updated (automatic IQ correction only).  Used by twinpeaks_kernel (configure.cu), not by the receiver kernels.
struct TwinPeaks { int state, counter, runs, restarts; float phase; };
__device__ __forceinline__ bool twinpeaks_active(const TwinPeaks &t) { return t.state == 2 || t.state == 0; }
__device__ __forceinline__ TwinPeaks twinpeaks_block(TwinPeaks t, float teta1, float teta3)
{
    if (t.state == 2) t.counter++;
    if (t.counter > 1000) { t.state = 0; t.counter = 0; t.phase = 0.0f; t.runs = 0; }
    if (teta3 != 0.0f && t.state == 0) {
        const float cur = asinf(__fdiv_rn(teta1, teta3));
        if (t.runs == 0) t.phase = cur;
        else t.phase = (float)(0.05 * (double)cur + 0.95 * (double)t.phase);
        t.runs++;
        if (t.runs == 50) {
            if ((double)fabsf(t.phase) > (3.14159265358979323846 / 8.0)) {
                t.state = 4;
                t.restarts++;
                if (t.restarts >= 4) { t.state = 3; t.restarts = 0; }
            } else { t.state = 1; t.restarts = 0; }
        }
    }
    return t;
}
__device__ __forceinline__ TwinPeaks twinpeaks_load(const ChanState *st) { TwinPeaks t = { st->tw_state, st->tw_counter, st->tw_runs, st->tw_restarts, st->tw_phase }; return t; }
__device__ __forceinline__ void twinpeaks_store(ChanState *st, const TwinPeaks &t)
{
    st->tw_state = t.state; st->tw_counter = t.counter; st->tw_runs = t.runs; st->tw_restarts = t.restarts; st->tw_phase = t.phase;
}

// float -> int32 -> << 16 output formatting, audio_driver.c:2911-2922.  In range this is C
// truncation toward zero; out-of-range values (undefined behaviour in the reference) saturate.
__device__ __forceinline__ int32_t format_audio_word(float a)
{
    return (int32_t)((uint32_t)__float2int_rz(a) << 16);
}

}  // namespace uhsdr
