// demod_device.cuh -- sample-serial demodulators (run by one lane per channel).  The state type is a template
// parameter: ChanState (general kernel, state staged in shared memory) or SerState (rx_serial.cu, per-thread copy).
#pragma once
#include "dsp_device.cuh"
#include "uhsdr_b200.h"

namespace uhsdr {

// AudioDriver_FadeLeveler, audio_driver.c:1911-1923
template <class S>
__device__ __forceinline__ float fade_leveler(const ChanParams &p, S &st, float audio, float corr)
{
    st.fade_dc27 = __fadd_rn(__fmul_rn(p.sam_mtauR, st.fade_dc27), __fmul_rn(p.sam_onem_mtauR, audio));
    st.fade_dc_insert = __fadd_rn(__fmul_rn(p.sam_mtauI, st.fade_dc_insert), __fmul_rn(p.sam_onem_mtauI, corr));
    return __fsub_rn(__fadd_rn(audio, st.fade_dc_insert), st.fade_dc27);
}

// AudioDriver_DemodSAM, audio_driver.c:1990-2166, for nb blocks of n_per_blk decimated samples.
// AM: envelope sqrt(i^2+q^2) (+ fade leveler).  SAM: NCO, phase detector, optional sideband
// selection through two 7-stage all-pass networks, 2nd-order loop filter.
template <class S>
__device__ inline void demod_am_sam(const ChanParams &p, S &st, const float *__restrict__ pool,
                                    const float *ib, const float *qb, float *a, int nb, int n_per_blk)
{
    const float sampleRate = (float)p.decimated_freq;
    if (p.mode == UHSDR_DEMOD_AM) {
        for (int i = 0; i < nb * n_per_blk; i++) {
            float audio = __fsqrt_rn(__fadd_rn(__fmul_rn(ib[i], ib[i]), __fmul_rn(qb[i], qb[i])));
            if (p.fade_leveler) audio = fade_leveler(p, st, audio, 0.0f);
            a[i] = audio;
        }
        return;
    }
    const float *c0 = pool + p.sam_c0, *c1 = pool + p.sam_c1;
    const double two_pi = 2.0 * (double)3.14159265358979f;
    for (int b = 0; b < nb; b++) {
        for (int n = 0; n < n_per_blk; n++) {
            const int i = b * n_per_blk + n;
            float Sin, Cos;
            sincosf(st.sam_phs, &Sin, &Cos);
            const float ai = __fmul_rn(Cos, ib[i]), bi = __fmul_rn(Sin, ib[i]);
            const float aq = __fmul_rn(Cos, qb[i]), bq = __fmul_rn(Sin, qb[i]);
            const float corr0 = __fadd_rn(ai, bq), corr1 = __fadd_rn(-bi, aq);
            float audio;
            if (p.sam_sideband != UHSDR_SAM_SIDEBAND_BOTH) {
                st.sam_a[0] = st.sam_dsI; st.sam_b[0] = bi; st.sam_c[0] = st.sam_dsQ; st.sam_d[0] = aq;
                st.sam_dsI = ai; st.sam_dsQ = bq;
                for (int j = 0; j < 7; j++) {
                    const int k = 3 * j;
                    const float k0 = __ldg(c0 + j), k1 = __ldg(c1 + j);
                    st.sam_a[k + 3] = __fadd_rn(__fmul_rn(k0, __fsub_rn(st.sam_a[k], st.sam_a[k + 5])), st.sam_a[k + 2]);
                    st.sam_b[k + 3] = __fadd_rn(__fmul_rn(k1, __fsub_rn(st.sam_b[k], st.sam_b[k + 5])), st.sam_b[k + 2]);
                    st.sam_c[k + 3] = __fadd_rn(__fmul_rn(k0, __fsub_rn(st.sam_c[k], st.sam_c[k + 5])), st.sam_c[k + 2]);
                    st.sam_d[k + 3] = __fadd_rn(__fmul_rn(k1, __fsub_rn(st.sam_d[k], st.sam_d[k + 5])), st.sam_d[k + 2]);
                }
                const float ai_ps = st.sam_a[21], bi_ps = st.sam_b[21], bq_ps = st.sam_c[21], aq_ps = st.sam_d[21];
                for (int j = 23; j > 0; j--) {
                    st.sam_a[j] = st.sam_a[j - 1]; st.sam_b[j] = st.sam_b[j - 1];
                    st.sam_c[j] = st.sam_c[j - 1]; st.sam_d[j] = st.sam_d[j - 1];
                }
                if (p.sam_sideband == UHSDR_SAM_SIDEBAND_LSB) audio = __fsub_rn(__fadd_rn(ai_ps, bi_ps), __fsub_rn(aq_ps, bq_ps));
                else audio = __fadd_rn(__fsub_rn(ai_ps, bi_ps), __fadd_rn(aq_ps, bq_ps));
            } else {
                audio = corr0;
            }
            if (p.fade_leveler) audio = fade_leveler(p, st, audio, corr0);
            a[i] = audio;

            const float phzerror = atan2f(corr1, corr0);
            const float del_out = st.sam_fil_out;
            st.sam_omega2 = __fadd_rn(st.sam_omega2, __fmul_rn(p.sam_g2, phzerror));
            if (st.sam_omega2 < p.sam_omega_min) st.sam_omega2 = p.sam_omega_min;
            else if (st.sam_omega2 > p.sam_omega_max) st.sam_omega2 = p.sam_omega_max;
            st.sam_fil_out = __fadd_rn(__fmul_rn(p.sam_g1, phzerror), st.sam_omega2);
            float phs = __fadd_rn(st.sam_phs, del_out);
            // wrap to [0, 2 pi): the comparisons and corrections are double expressions (:2146-2147)
            while ((double)phs >= two_pi) phs = (float)((double)phs - two_pi);
            while (phs < 0.0f) phs = (float)((double)phs + two_pi);
            st.sam_phs = phs;
        }
        // carrier-offset display value, once per call of the reference function (:2150-2162)
        st.sam_count++;
        if (st.sam_count > 50) {
            float carrier = (float)(0.1 * (double)__fmul_rn(st.sam_omega2, sampleRate) / two_pi);
            carrier = (float)((double)carrier + 0.9 * (double)st.sam_lowpass);
            st.carrier_freq_offset = (int)carrier;
            st.sam_count = 0;
            st.sam_lowpass = carrier;
        }
    }
}

// AudioDriver_DemodFM, audio_driver.c:1544-1737 (with the 3 x Goertzel subaudible-tone detector :1665-1734), nb blocks at
// 48 ksps.  Returns a bit mask: bit b set = block b un-squelched (signal_active).
template <class S>
__device__ inline int demod_fm(const ChanParams &p, S &st, const float *__restrict__ pool,
                               const float *ib, const float *qb, float *a, int nb)
{
    int mask = 0;
    for (int b = 0; b < nb; b++) {
        if (p.fm_translate_on) {
            float first_hp = 0.0f;
            for (int n = 0; n < BLK; n++) {
                const int i = b * BLK + n;
                const float y = __fsub_rn(__fmul_rn(st.fm_i_prev, qb[i]), __fmul_rn(ib[i], st.fm_q_prev));
                const float x = __fadd_rn(__fmul_rn(st.fm_i_prev, ib[i]), __fmul_rn(qb[i], st.fm_q_prev));
                const float angle = atan2f(y, x);
                // squelch noise high-pass (6-stage lattice) on the raw angle (:1594); only the
                // block's first output sample is used (:1597-1598), but the filter runs on all
                const float hp = lattice_step(angle, st.sql_s, pool + p.sql.k_off, pool + p.sql.v_off, p.sql.n);
                if (n == 0) first_hp = hp;
                // de-emphasis: a = lpf_prev + 0.05*(angle - lpf_prev), double expression (:1566)
                const float av = (float)((double)st.fm_lpf_prev + (0.05 * (double)__fsub_rn(angle, st.fm_lpf_prev)));
                st.fm_lpf_prev = av;
                if (p.fm_tone_det) {                 // AudioFilter_GoertzelInput x 3 on the de-emphasised audio (:1684-1692)
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const float b0 = __fadd_rn(__fsub_rn(__fmul_rn(p.fm_gz_r[k], st.fm_gz[3 * k + 1]), st.fm_gz[3 * k + 2]), av);
                        st.fm_gz[3 * k + 2] = st.fm_gz[3 * k + 1]; st.fm_gz[3 * k + 1] = b0; st.fm_gz[3 * k] = b0;
                    }
                }
                // audio gate :1571-1587: open when un-squelched (no tone detection), or when the tone is there, or squelch off
                if ((!st.fm_squelched && !p.fm_tone_det) || (st.fm_tone_detected && p.fm_tone_det) || !p.fm_sql_threshold) {
                    const float hb = (float)(0.96 * (double)__fsub_rn(__fadd_rn(st.fm_hpf_prev_b, av), st.fm_hpf_prev_a));
                    st.fm_hpf_prev_a = av;
                    st.fm_hpf_prev_b = hb;
                    a[i] = hb;
                } else {
                    a[i] = 0.0f;
                }
                st.fm_q_prev = qb[i];
                st.fm_i_prev = ib[i];
            }
            st.fm_sql_avg = (float)(((double)(1 - 0.005) * (double)st.fm_sql_avg) + (0.005 * (double)__fsqrt_rn(fabsf(first_hp))));
            st.fm_count = (st.fm_count + 1) % 200;
            if (st.fm_count == 0) {
                if ((double)st.fm_sql_avg > 0.175) st.fm_sql_avg = (float)0.175;
                float s = __fmul_rn(st.fm_sql_avg, 172.0f);
                if (s > 24.0f) s = 24.0f;
                s = __fsub_rn(22.0f, s);
                const int thr = p.fm_sql_threshold;
                if (thr == 0) st.fm_squelched = 0;
                else if (st.fm_squelched) { if (s >= (float)(thr + 3)) st.fm_squelched = 0; }
                else if (thr > 3) { if (s < (float)(thr - 3)) st.fm_squelched = 1; }
                else { if (s < (float)thr) st.fm_squelched = 1; }
            }
            if (p.fm_tone_det) {
                // every 400 blocks: ratio of the on-frequency energy to the mean of the two off-frequency ones, smoothed,
                // thresholded at 1.75 and debounced (:1694-1729)
                st.fm_gcount++;
                if (st.fm_gcount >= 400) {
                    float en[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {        // AudioFilter_GoertzelEnergy, audio_filter.c:1297-1305
                        const float ea = __fsub_rn(st.fm_gz[3 * k + 1], __fmul_rn(st.fm_gz[3 * k + 2], p.fm_gz_cos[k]));
                        const float eb = __fmul_rn(st.fm_gz[3 * k + 2], p.fm_gz_sin[k]);
                        st.fm_gz[3 * k] = 0.0f; st.fm_gz[3 * k + 1] = 0.0f; st.fm_gz[3 * k + 2] = 0.0f;
                        en[k] = __fsqrt_rn(__fadd_rn(__fmul_rn(ea, ea), __fmul_rn(eb, eb)));
                    }
                    const float s = __fadd_rn(en[0], en[1]), r = en[2];
                    st.fm_subdet = (float)(((1 - 0.9) * (double)st.fm_subdet) + ((double)__fdiv_rn(r, __fdiv_rn(s, 2.0f)) * 0.9));
                    if ((double)st.fm_subdet > 1.75) { st.fm_tdet++; if (st.fm_tdet > 5) st.fm_tdet = 5; }
                    else if (st.fm_tdet) st.fm_tdet--;
                    st.fm_tone_detected = st.fm_tdet >= 2 ? 1 : 0;
                    st.fm_gcount = 0;
                }
            } else {
                st.fm_tone_detected = 1;            // detection disabled: always "detected" (:1731-1734)
            }
        }
        if (!st.fm_squelched) mask |= (1 << b);
    }
    return mask;
}

}  // namespace uhsdr
