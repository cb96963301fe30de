// rx_serial2.cu -- sample-serial half of the split general receiver path, second generation (shipping build; the exact build
// keeps rx_serial_kernel and its reference-order arithmetic).  One channel per thread as before, but the per-sample cost of the
// recurrences is cut to what the tensor-core kernel's serial roles pay (rx_ssb_tc.cu):
//   * fused multiply-adds; biquads in the form whose sample-to-sample critical path is one FMA per stage; identity stages skipped;
//   * the WDSP AGC (audio_agc.c:349-595) without its rescan loop: the sliding maximum over attack_buffsize = 6 ND + 1 samples is
//     max(prefix maximum of the block, block maxima of the five blocks before, suffix maximum of the sixth) -- the same exact value
//     the reference finds by rescanning its ring -- with the per-block suffix maxima and the sample ring in SHARED memory,
//     [slot][lane], instead of a per-thread local-memory ring; gain law with the bit-level log10 and a fast division;
//   * every per-block array in registers (block sizes are template parameters), scratch rows fetched one block ahead with 16-byte
//     loads, 32-byte output stores.
// Chain after the FIR stages (audio_driver.c): AM / SAM demodulation (:1990-2166) | LMS notch (:1746-1763) | lattice pre-filter
// (:2473-2475) | AGC (:2485) | spectral NR hand-off (phases, :2501-2509) | gain + 4-stage biquad (:2513-2527) | polyphase
// interpolator (:2560-2577) | anti-alias lattice (:2581-2583) | treble biquad (:2832) | formatting (:2845-2941).  FM (:1544-1737) runs
// its discriminator, de-emphasis and squelch per sample and needs none of the above but the treble stage.
#include <type_traits>

#include "demod_device.cuh"
#include "kernels.h"

#if !UHSDR_EXACT

namespace uhsdr {

namespace {

constexpr int S2_THREADS = 32;
constexpr int NSUF = 6;                 // blocks whose suffix maxima are kept: attack_buffsize - 1 = 6 ND

// AM / SAM demodulator (AudioDriver_DemodSAM, audio_driver.c:1990-2166; demod_am_sam in demod_device.cuh is the operation-for-
// operation form the other kernels run): constants and the PLL / fade-leveler state in registers, the four 7-stage all-pass chains of
// the sideband selector in shared memory.  The reference keeps each chain as a 24-element array shifted by one per sample
// (a[3 s], a[3 s + 1], a[3 s + 2] = signal s now, one and two samples ago; signal 0 = the chain's input, signal s = the output of
// section s; a[3 s + 3] = k (a[3 s] - a[3 s + 5]) + a[3 s + 2]): that is a cascade of second-order all-pass sections whose state is
// the last two values of its eight signals.  Here: AP[chain][signal][slot][lane], slot = sample parity, the value of two samples
// ago is read and then overwritten by the new one.  Same operations in the same order as the reference, so the same bits.
struct Demod2 {
    bool am, sel, lsb, fade;
    float mtauR, onem_mtauR, mtauI, onem_mtauI, g1, g2, omin, omax, rate;
    float k0[7], k1[7];
    // state
    float fil_out, lowpass, omega2, phs, dsI, dsQ, dc27, dc_insert;
    int count, carrier, par;
};
constexpr int AP_FLOATS = 4 * 8 * 2 * S2_THREADS;       // all-pass state per CTA

__device__ __forceinline__ void demod2_load(Demod2 &d, const ChanParams &p, const ChanState &g, const float *__restrict__ pool, float *ap)
{
    d.am = p.mode == UHSDR_DEMOD_AM; d.sel = !d.am && p.sam_sideband != UHSDR_SAM_SIDEBAND_BOTH; d.lsb = p.sam_sideband == UHSDR_SAM_SIDEBAND_LSB;
    d.fade = p.fade_leveler != 0;
    d.mtauR = p.sam_mtauR; d.onem_mtauR = p.sam_onem_mtauR; d.mtauI = p.sam_mtauI; d.onem_mtauI = p.sam_onem_mtauI;
    d.g1 = p.sam_g1; d.g2 = p.sam_g2; d.omin = p.sam_omega_min; d.omax = p.sam_omega_max; d.rate = (float)p.decimated_freq;
#pragma unroll
    for (int j = 0; j < 7; j++) { d.k0[j] = d.am ? 0.0f : __ldg(pool + p.sam_c0 + j); d.k1[j] = d.am ? 0.0f : __ldg(pool + p.sam_c1 + j); }
    d.fil_out = g.sam_fil_out; d.lowpass = g.sam_lowpass; d.omega2 = g.sam_omega2; d.phs = g.sam_phs; d.dsI = g.sam_dsI; d.dsQ = g.sam_dsQ;
    d.dc27 = g.fade_dc27; d.dc_insert = g.fade_dc_insert; d.count = g.sam_count; d.carrier = g.carrier_freq_offset; d.par = 0;
    if (d.sel) {
        for (int sg = 0; sg < 8; sg++) {
            // slot 0 = two samples ago (a[3 s + 2]), slot 1 = one sample ago (a[3 s + 1])
            ap[((0 * 8 + sg) * 2 + 0) * S2_THREADS] = g.sam_a[3 * sg + 2]; ap[((0 * 8 + sg) * 2 + 1) * S2_THREADS] = g.sam_a[3 * sg + 1];
            ap[((1 * 8 + sg) * 2 + 0) * S2_THREADS] = g.sam_b[3 * sg + 2]; ap[((1 * 8 + sg) * 2 + 1) * S2_THREADS] = g.sam_b[3 * sg + 1];
            ap[((2 * 8 + sg) * 2 + 0) * S2_THREADS] = g.sam_c[3 * sg + 2]; ap[((2 * 8 + sg) * 2 + 1) * S2_THREADS] = g.sam_c[3 * sg + 1];
            ap[((3 * 8 + sg) * 2 + 0) * S2_THREADS] = g.sam_d[3 * sg + 2]; ap[((3 * 8 + sg) * 2 + 1) * S2_THREADS] = g.sam_d[3 * sg + 1];
        }
    }
}

__device__ __forceinline__ void demod2_store(ChanState &g, const Demod2 &d, const float *ap)
{
    g.sam_fil_out = d.fil_out; g.sam_lowpass = d.lowpass; g.sam_omega2 = d.omega2; g.sam_phs = d.phs; g.sam_dsI = d.dsI; g.sam_dsQ = d.dsQ;
    g.fade_dc27 = d.dc27; g.fade_dc_insert = d.dc_insert; g.sam_count = d.count; g.carrier_freq_offset = d.carrier;
    if (d.sel) {
        // slot par = two samples before the next one = a[3 s + 2] after the reference's shift; the other slot = a[3 s + 1]; a[3 s]
        // is written before it is read (a[0] from dsI / the input, a[3 s] by section s - 1) and carries no state
        const int o = d.par, n = d.par ^ 1;
        for (int sg = 0; sg < 8; sg++) {
            g.sam_a[3 * sg + 2] = ap[((0 * 8 + sg) * 2 + o) * S2_THREADS]; g.sam_a[3 * sg + 1] = ap[((0 * 8 + sg) * 2 + n) * S2_THREADS];
            g.sam_b[3 * sg + 2] = ap[((1 * 8 + sg) * 2 + o) * S2_THREADS]; g.sam_b[3 * sg + 1] = ap[((1 * 8 + sg) * 2 + n) * S2_THREADS];
            g.sam_c[3 * sg + 2] = ap[((2 * 8 + sg) * 2 + o) * S2_THREADS]; g.sam_c[3 * sg + 1] = ap[((2 * 8 + sg) * 2 + n) * S2_THREADS];
            g.sam_d[3 * sg + 2] = ap[((3 * 8 + sg) * 2 + o) * S2_THREADS]; g.sam_d[3 * sg + 1] = ap[((3 * 8 + sg) * 2 + n) * S2_THREADS];
            g.sam_a[3 * sg] = sg ? g.sam_a[3 * sg - 1] : 0.0f; g.sam_b[3 * sg] = sg ? g.sam_b[3 * sg - 1] : 0.0f;
            g.sam_c[3 * sg] = sg ? g.sam_c[3 * sg - 1] : 0.0f; g.sam_d[3 * sg] = sg ? g.sam_d[3 * sg - 1] : 0.0f;
        }
    }
}

// one all-pass chain, one sample: x = the chain's input now; returns the output of section 7 (the reference's a[21])
__device__ __forceinline__ float allpass7(float *chain, int par, const float (&k)[7], float x)
{
    float m2[8];
#pragma unroll
    for (int sg = 0; sg < 8; sg++) m2[sg] = chain[(sg * 2 + par) * S2_THREADS];       // every signal two samples ago
#pragma unroll
    for (int j = 0; j < 7; j++) {
        chain[(j * 2 + par) * S2_THREADS] = x;
        x = __fadd_rn(__fmul_rn(k[j], __fsub_rn(x, m2[j + 1])), m2[j]);
    }
    chain[(7 * 2 + par) * S2_THREADS] = x;
    return x;
}

// AudioDriver_FadeLeveler, audio_driver.c:1911-1923
__device__ __forceinline__ float fade2(Demod2 &d, float audio, float corr)
{
    d.dc27 = __fadd_rn(__fmul_rn(d.mtauR, d.dc27), __fmul_rn(d.onem_mtauR, audio));
    d.dc_insert = __fadd_rn(__fmul_rn(d.mtauI, d.dc_insert), __fmul_rn(d.onem_mtauI, corr));
    return __fsub_rn(__fadd_rn(audio, d.dc_insert), d.dc27);
}

template <int ND>
__device__ __forceinline__ void demod2_block(Demod2 &d, float *ap, const float (&ib)[ND], const float (&qb)[ND], float (&a)[ND])
{
    if (d.am) {
#pragma unroll
        for (int i = 0; i < ND; i++) {
            float audio = __fsqrt_rn(__fadd_rn(__fmul_rn(ib[i], ib[i]), __fmul_rn(qb[i], qb[i])));
            if (d.fade) audio = fade2(d, audio, 0.0f);
            a[i] = audio;
        }
        return;
    }
    const double two_pi = 2.0 * (double)3.14159265358979f;
#pragma unroll 1
    for (int i = 0; i < ND; i++) {
        float Sin, Cos;
        sincosf(d.phs, &Sin, &Cos);
        const float ai = __fmul_rn(Cos, ib[i]), bi = __fmul_rn(Sin, ib[i]);
        const float aq = __fmul_rn(Cos, qb[i]), bq = __fmul_rn(Sin, qb[i]);
        const float corr0 = __fadd_rn(ai, bq), corr1 = __fadd_rn(-bi, aq);
        float audio = corr0;
        if (d.sel) {
            const float ai_ps = allpass7(ap + 0 * 16 * S2_THREADS, d.par, d.k0, d.dsI);
            const float bi_ps = allpass7(ap + 1 * 16 * S2_THREADS, d.par, d.k1, bi);
            const float bq_ps = allpass7(ap + 2 * 16 * S2_THREADS, d.par, d.k0, d.dsQ);
            const float aq_ps = allpass7(ap + 3 * 16 * S2_THREADS, d.par, d.k1, aq);
            d.dsI = ai; d.dsQ = bq;
            d.par ^= 1;
            if (d.lsb) audio = __fsub_rn(__fadd_rn(ai_ps, bi_ps), __fsub_rn(aq_ps, bq_ps));
            else audio = __fadd_rn(__fsub_rn(ai_ps, bi_ps), __fadd_rn(aq_ps, bq_ps));
        }
        if (d.fade) audio = fade2(d, audio, corr0);
        a[i] = audio;
        const float phzerror = atan2f(corr1, corr0);
        const float del_out = d.fil_out;
        d.omega2 = __fadd_rn(d.omega2, __fmul_rn(d.g2, phzerror));
        if (d.omega2 < d.omin) d.omega2 = d.omin;
        else if (d.omega2 > d.omax) d.omega2 = d.omax;
        d.fil_out = __fadd_rn(__fmul_rn(d.g1, phzerror), d.omega2);
        float phs = __fadd_rn(d.phs, del_out);
        // wrap to [0, 2 pi): the comparisons and corrections are double expressions (:2146-2147)
        while ((double)phs >= two_pi) phs = (float)((double)phs - two_pi);
        while (phs < 0.0f) phs = (float)((double)phs + two_pi);
        d.phs = phs;
    }
    // carrier-offset display value, once per call of the reference function (:2150-2162)
    d.count++;
    if (d.count > 50) {
        float carrier = (float)(0.1 * (double)__fmul_rn(d.omega2, d.rate) / two_pi);
        carrier = (float)((double)carrier + 0.9 * (double)d.lowpass);
        d.carrier = (int)carrier;
        d.count = 0;
        d.lowpass = carrier;
    }
}

// arm_fir_interpolate_f32 (:2560-2577) for one block: output n = i L + j uses taps c[(L-1-j) + k L] on ip[INTERP_HIST - (P-1) + i + k]
template <int L, int P, int ND>
__device__ __forceinline__ void interp_block2(const float (&ip)[INTERP_HIST + ND], const float *__restrict__ ic, float (&o48)[BLK])
{
    float c[L][P];
#pragma unroll
    for (int j = 0; j < L; j++)
#pragma unroll
        for (int k = 0; k < P; k++) c[j][k] = __ldg(ic + (L - 1 - j) + k * L);
#pragma unroll
    for (int i = 0; i < ND; i++)
#pragma unroll
        for (int j = 0; j < L; j++) {
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < P; k++) sum = fmaf(ip[INTERP_HIST - (P - 1) + i + k], c[j][k], sum);
            o48[i * L + j] = sum;
        }
}

// LMS automatic notch on one decimated block (AudioDriver_NotchFilter around arm_lms_norm_f32), state in global memory: the
// filter is rare and 64 taps long; same arithmetic as rx_serial.cu
template <int ND>
__device__ __noinline__ void notch_block2(ChanState &st, float (&buf)[ND], float mu)
{
    for (int i = 0; i < ND; i++) st.notch_delay[st.notch_inbuf + i] = buf[i];
    float energy = st.notch_energy, x0 = st.notch_x0;
    int head = st.notch_head;
    for (int i = 0; i < ND; i++) {
        const float in = buf[i];
        st.notch_x[(head + NOTCH_TAPS - 1) & (NOTCH_TAPS - 1)] = in;
        energy = __fsub_rn(energy, __fmul_rn(x0, x0));
        energy = __fadd_rn(energy, __fmul_rn(in, in));
        float sum = 0.0f;
        for (int k = 0; k < NOTCH_TAPS; k++) sum = __fadd_rn(sum, __fmul_rn(st.notch_x[(head + k) & (NOTCH_TAPS - 1)], st.notch_coef[k]));
        const float d = st.notch_delay[st.notch_outbuf + i];
        const float e = __fsub_rn(d, sum);
        buf[i] = e;
        const float w = __fdiv_rn(__fmul_rn(e, mu), __fadd_rn(energy, 0.000000119209289f));
        for (int k = 0; k < NOTCH_TAPS; k++)
            st.notch_coef[k] = __fadd_rn(st.notch_coef[k], __fmul_rn(w, st.notch_x[(head + k) & (NOTCH_TAPS - 1)]));
        x0 = st.notch_x[head];
        head = (head + 1) & (NOTCH_TAPS - 1);
    }
    st.notch_energy = energy; st.notch_x0 = x0; st.notch_head = head;
    st.notch_inbuf += ND;
    st.notch_outbuf = st.notch_inbuf + ND;
    st.notch_inbuf %= NOTCH_DELAY;
    st.notch_outbuf %= NOTCH_DELAY;
}

// output stage shared by all variants: anti-alias lattice (6 stages where present), treble biquad, x10, formatting, 32-byte stores
struct OutStage {
    float ak[6], av[7], as_[6];
    float tc[5], tt;
    BiquadS ts;
    bool aa_on, tr_unity;
};

__device__ __forceinline__ void out_init(OutStage &o, const ChanParams &p, const ChanState &st, const float *__restrict__ pool, bool fm)
{
    o.aa_on = p.aa.n == 6 && !fm;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        o.ak[j] = o.aa_on ? __ldg(pool + p.aa.k_off + j) : 0.0f;
        o.av[j] = o.aa_on ? __ldg(pool + p.aa.v_off + j) : 0.0f;
        o.as_[j] = o.aa_on ? st.aa_s[j] : 0.0f;
    }
    o.av[6] = o.aa_on ? __ldg(pool + p.aa.v_off + 6) : 1.0f;
#pragma unroll
    for (int q = 0; q < 5; q++) o.tc[q] = p.bq2[q];
    o.ts = st.bq2;
    o.tt = fmaf(o.tc[3], o.ts.y1, fmaf(o.tc[1], o.ts.x1, fmaf(o.tc[2], o.ts.x2, __fmul_rn(o.tc[4], o.ts.y2))));
    // a 0 dB shelf with consistent state is the identity (see rx_ssb_tc.cu): not computed, the state follows the signal
    o.tr_unity = o.tc[0] == 1.0f && o.tc[1] == -o.tc[3] && o.tc[2] == -o.tc[4] && o.ts.x1 == o.ts.y1 && o.ts.x2 == o.ts.y2;
}

__device__ __forceinline__ void out_block(OutStage &o, const float (&o48)[BLK], bool muted, int2 *__restrict__ dst, float *__restrict__ dst_f)
{
#pragma unroll
    for (int n = 0; n < BLK; n += 4) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float y = o48[n + e];
            if (o.aa_on) {
                float f = y, acc = 0.0f, fn = y;
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    const float gg = o.as_[q];
                    fn = fmaf(-o.ak[q], gg, f);
                    const float gn = fmaf(fn, o.ak[q], gg);
                    acc = fmaf(gn, o.av[q], acc);
                    if (q > 0) o.as_[q - 1] = gn;
                    f = fn;
                }
                y = fmaf(fn, o.av[6], acc);
                o.as_[5] = fn;
            }
            float z = y;
            if (!o.tr_unity) {
                const float w = fmaf(o.tc[2], o.ts.x1, __fmul_rn(o.tc[4], o.ts.y1));
                z = fmaf(o.tc[0], y, o.tt);
                o.tt = fmaf(o.tc[3], z, fmaf(o.tc[1], y, w));
                o.ts.x2 = o.ts.x1; o.ts.x1 = y; o.ts.y2 = o.ts.y1; o.ts.y1 = z;
            } else {
                o.ts.x2 = o.ts.x1; o.ts.x1 = y; o.ts.y2 = o.ts.x2; o.ts.y1 = y;
            }
            v[e] = muted ? 0.0f : __fmul_rn(z, 10.0f);                   // LINE_OUT_SCALING_FACTOR (:2860)
        }
        const int w0 = muted ? 0 : format_audio_word(v[0]), w1 = muted ? 0 : format_audio_word(v[1]);
        const int w2 = muted ? 0 : format_audio_word(v[2]), w3 = muted ? 0 : format_audio_word(v[3]);
        asm volatile("st.global.v8.b32 [%0], {%1, %1, %2, %2, %3, %3, %4, %4};" ::"l"(dst + n), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        if (dst_f) *reinterpret_cast<float4 *>(dst_f + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__device__ __forceinline__ void out_save(const OutStage &o, ChanState &st)
{
    if (o.aa_on) {
#pragma unroll
        for (int j = 0; j < 6; j++) st.aa_s[j] = o.as_[j];
    }
    st.bq2 = o.ts;
}

// ---- SSB / AM / SAM: ND decimated samples per block (8 at 12 ksps, 16 at 24 ksps) ----
// EX: the stages in front of the AGC's DC remover in the reference's arithmetic (separate multiply and add, frexpf-based log10, IEEE
// division).  AM / SAM need it: their DC remover (audio_agc.c:577-594, pole 0.9999) holds a state ~1e4 x the carrier level whose
// rounding -- and with it the output at the 1e-4 level -- follows every last bit of its input (see rx_generic.cu fir_dot_exact).
template <int ND, bool EX>
__device__ void serial2_dec(const RxArgs &a, int phase, int slot, int ch, float *smem)
{
    constexpr int L = BLK / ND;
    const int lane = threadIdx.x;
    const ChanParams &p = a.params[ch];
    ChanState &st = a.state[ch];
    const float *__restrict__ pool = a.pool;
    const bool amsam = p.topo == TOPO_AM_SAM;
    const bool notch = p.notch_enable != 0;
    float *ring = smem + lane;                                   // AGC sample ring [AGC_RB][32]
    float *suf = smem + AGC_RB * S2_THREADS + lane;              // suffix maxima [NSUF][ND][32]
    float *sc = a.scratch + (size_t)slot * (size_t)a.scratch_stride;
    const size_t half = (size_t)a.nblocks * ND;
    const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
    int2 *__restrict__ audio = reinterpret_cast<int2 *>(a.audio) + chan_base;
    float *__restrict__ audio_f = a.audio_f ? a.audio_f + chan_base : nullptr;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;

    Demod2 ds;
    float *apc = smem + (AGC_RB + NSUF * 16) * S2_THREADS + lane;  // all-pass chains of the SAM sideband selector
    if (amsam && phase != 2) demod2_load(ds, p, st, pool, apc);
    // lattice pre-filter, front-padded to 10 stages (k = v = 0 stages pass the sample through)
    float pk[10], pv[11], ps[10];
    const int pn = p.pre.n, ppad = 10 - pn;
#pragma unroll
    for (int j = 0; j < 10; j++) {
        pk[j] = (j >= ppad) ? __ldg(pool + p.pre.k_off + (j - ppad)) : 0.0f;
        pv[j] = (j >= ppad) ? __ldg(pool + p.pre.v_off + (j - ppad)) : 0.0f;
        ps[j] = (j >= ppad) ? st.pre_s[j - ppad] : 0.0f;
    }
    pv[10] = (pn > 0) ? __ldg(pool + p.pre.v_off + pn) : 1.0f;
    // AGC
    const AgcP ap = p.agc;
    const bool agc_off = ap.mode == 5;
    AgcRun ar = { st.agc_out_index, st.agc_in_index, st.agc_ring_max, st.agc_volts, st.agc_save_volts, st.agc_fast_backaverage,
                  st.agc_hang_backaverage, st.agc_hang_counter, st.agc_decay_type, st.agc_state, st.agc_action, st.agc_hang_action };
    const bool remove_dc = ap.remove_dc && !agc_off;
    float agc_wold = st.agc_wold;
    int head = 0;                                                // slot of the oldest kept block (block -6)
    float m5 = 0.0f;                                             // maximum over the blocks -1 .. -5
    if (phase != 2 && !agc_off) {
        for (int i = 0; i < AGC_RB; i++) ring[i * S2_THREADS] = st.agc_ring[i];
        // suffix maxima of the six blocks in front of this launch, from the ring (newest sample at in_index)
        for (int b = 0; b < NSUF; b++) {                          // b = 0: block -6 ... b = 5: block -1
            float m = 0.0f;
            for (int j = ND - 1; j >= 0; j--) {
                int idx = ar.in_index - ((NSUF - 1 - b) * ND + (ND - 1 - j));
                idx %= AGC_RB; if (idx < 0) idx += AGC_RB;
                m = fmaxf(m, fabsf(ring[idx * S2_THREADS]));
                suf[(b * ND + j) * S2_THREADS] = m;
            }
        }
        for (int b = 1; b < NSUF; b++) m5 = fmaxf(m5, suf[(b * ND) * S2_THREADS]);
    }
    // (phase 2 does not run the AGC and must not look at its state: phase 1 of the next time slice may be writing it)
    const bool any_hang = phase != 2 && __any_sync(__activemask(), ap.hang_enable || ar.state == 2 || ar.state == 4 || ar.decay_type != 0 || ar.hang_counter > 0);
    // biquad cascade: t-form (see rx_ssb_tc.cu), identity stages skipped per thread
    float bc[4][5], tq[4];
    BiquadS bs[4];
    bool ident[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
#pragma unroll
        for (int q = 0; q < 5; q++) bc[s][q] = p.bq1[s][q];
        bs[s] = st.bq1[s];
        ident[s] = bc[s][0] == 1.0f && bc[s][1] == 0.0f && bc[s][2] == 0.0f && bc[s][3] == 0.0f && bc[s][4] == 0.0f;
        tq[s] = fmaf(bc[s][3], bs[s].y1, fmaf(bc[s][1], bs[s].x1, fmaf(bc[s][2], bs[s].x2, __fmul_rn(bc[s][4], bs[s].y2))));
    }
    const float scale_gain = p.scale_gain, notch_mu = p.notch_mu;
    float ip[INTERP_HIST + ND];
#pragma unroll
    for (int i = 0; i < INTERP_HIST; i++) ip[i] = st.interp_hist[i];
    const int P = p.interp_plen;
    const float *__restrict__ ic = pool + p.interp_c;
    OutStage os;
    out_init(os, p, st, pool, false);

    // input of the first block
    float nxa[ND], nxb[ND];
    // volatile asm: the compiler must not sink these loads down to their first use one block later
    auto ld4 = [](float *d, const float *ptr) {
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(ptr));
    };
    auto fetch = [&](int blk) {
#pragma unroll
        for (int q = 0; q < ND / 4; q++) ld4(nxa + 4 * q, sc + (size_t)blk * ND + 4 * q);
        if (amsam && phase != 2) {
#pragma unroll
            for (int q = 0; q < ND / 4; q++) ld4(nxb + 4 * q, sc + half + (size_t)blk * ND + 4 * q);
        }
    };
    fetch(0);

    // LEAN (decided once per warp): no hang AGC, only the bass shelf of the biquad cascade active (the default plan), no
    // anti-alias lattice, unity treble shelf -- the common case gets a loop without those branches and state shuffles
    // (about 40 % fewer instructions per block, and a hot loop that fits the instruction cache); everything else the general loop.
    unsigned skipmask = 0;
#pragma unroll
    for (int s_ = 0; s_ < 4; s_++) if (ident[s_]) skipmask |= 1u << s_;
    const bool lean = __all_sync(__activemask(), !any_hang && skipmask == 0xbu && !os.aa_on && os.tr_unity && !notch);
    float xl1 = 0.0f, xl2 = 0.0f, ol1 = os.ts.x1, ol2 = os.ts.x2;
    auto run = [&](auto leanc) {
    constexpr bool LEAN = decltype(leanc)::value;
    for (int blk = 0; blk < a.nblocks; blk++) {
        float ad[ND], bq_[ND];
#pragma unroll
        for (int i = 0; i < ND; i++) { ad[i] = nxa[i]; bq_[i] = nxb[i]; }
        if (blk + 1 < a.nblocks) fetch(blk + 1);
        if (phase != 2) {
            if (amsam) {
                float bi[ND];
#pragma unroll
                for (int i = 0; i < ND; i++) bi[i] = ad[i];
                demod2_block<ND>(ds, apc, bi, bq_, ad);
            }
            if (notch) {
                // (the LMS routine is a real call taking the block by reference: give it a copy, so that `ad` itself never has its
                // address taken and stays in registers on the common path)
                float nb_[ND];
#pragma unroll
                for (int i = 0; i < ND; i++) nb_[i] = ad[i];
                notch_block2<ND>(st, nb_, notch_mu);
#pragma unroll
                for (int i = 0; i < ND; i++) ad[i] = nb_[i];
            }
            // lattice pre-filter (the per-sample loops are unrolled: rolled, the per-block arrays go to local memory and the kernel
            // runs twice as long)
#pragma unroll
            for (int i = 0; i < ND; i++) {
                float f = ad[i], acc = 0.0f, fn = f;
#pragma unroll
                for (int j = 0; j < 10; j++) {
                    const float gg = ps[j];
                    fn = EX ? __fsub_rn(f, __fmul_rn(pk[j], gg)) : fmaf(-pk[j], gg, f);
                    const float gn = EX ? __fadd_rn(__fmul_rn(fn, pk[j]), gg) : fmaf(fn, pk[j], gg);
                    acc = EX ? __fadd_rn(acc, __fmul_rn(gn, pv[j])) : fmaf(gn, pv[j], acc);
                    if (j > 0) ps[j - 1] = gn;
                    f = fn;
                }
                ad[i] = EX ? __fadd_rn(acc, __fmul_rn(fn, pv[10])) : fmaf(fn, pv[10], acc);
                ps[9] = fn;
            }
            if (agc_off) {
#pragma unroll
                for (int i = 0; i < ND; i++) ad[i] = __fmul_rn(ad[i], ap.fixed_gain);
            } else {
                // ---- WDSP AGC over the block ----
                const float *sf6 = suf + (head * ND) * S2_THREADS;            // suffix maxima of block -6
                float pmax = 0.0f, sm[ND], nsuf = 0.0f;
#pragma unroll
                for (int i = 0; i < ND; i++) {
                    sm[i] = sf6[i * S2_THREADS];
                    if (++ar.out_index >= AGC_RB) ar.out_index -= AGC_RB;
                    if (++ar.in_index >= AGC_RB) ar.in_index -= AGC_RB;
                    const float x = ad[i];
                    const float out_sample = ring[ar.out_index * S2_THREADS];
                    ring[ar.in_index * S2_THREADS] = x;
                    const float abs_out = fabsf(out_sample);
                    pmax = fmaxf(pmax, fabsf(x));
                    if (EX) {
                        ar.fast_backaverage = __fadd_rn(__fmul_rn(ap.fast_backmult, abs_out), __fmul_rn(ap.onemfast_backmult, ar.fast_backaverage));
                        ar.hang_backaverage = __fadd_rn(__fmul_rn(ap.hang_backmult, abs_out), __fmul_rn(ap.onemhang_backmult, ar.hang_backaverage));
                    } else {
                        ar.fast_backaverage = fmaf(ap.fast_backmult, abs_out, __fmul_rn(ap.onemfast_backmult, ar.fast_backaverage));
                        ar.hang_backaverage = fmaf(ap.hang_backmult, abs_out, __fmul_rn(ap.onemhang_backmult, ar.hang_backaverage));
                    }
                    ar.ring_max = fmaxf(pmax, fmaxf(m5, sm[i]));
                    const float dv = __fsub_rn(ar.ring_max, ar.volts);
                    const bool attack = ar.ring_max >= ar.volts;
                    float mult_sel = ap.attack_mult;
                    bool upd = true;
                    int nstate = ar.state;
                    if (LEAN || !any_hang) {
                        const int c0 = ar.volts > __fmul_rn(ap.pop_ratio, ar.fast_backaverage), c1 = ar.volts > ar.save_volts;
                        const bool fast = (((ar.state == 0) & c0) | ((ar.state == 1) & c1)) != 0;
                        ar.save_volts = (attack & (ar.state >= 2)) ? ar.volts : ar.save_volts;
                        mult_sel = attack ? ap.attack_mult : (fast ? ap.fast_decay_mult : ap.decay_mult);
                        nstate = attack ? 0 : (fast ? 1 : 3);
                    } else {
                        if (ar.hang_counter > 0) --ar.hang_counter;
                        if (attack) {
                            if (ar.state >= 2) ar.save_volts = ar.volts;
                            nstate = 0;
                        } else {
                            switch (ar.state) {
                            case 0:
                                if (ar.volts > __fmul_rn(ap.pop_ratio, ar.fast_backaverage)) { nstate = 1; mult_sel = ap.fast_decay_mult; }
                                else if (ap.hang_enable && (ar.hang_backaverage > ap.hang_level)) {
                                    nstate = 2; ar.hang_counter = (int)__fmul_rn(ap.hangtime, ap.sample_rate); ar.decay_type = 1; upd = false;
                                } else { nstate = 3; mult_sel = ap.decay_mult; ar.decay_type = 0; }
                                break;
                            case 1:
                                if (ar.volts > ar.save_volts) mult_sel = ap.fast_decay_mult;
                                else if (ar.hang_counter > 0) { nstate = 2; upd = false; }
                                else if (ar.decay_type == 0) { nstate = 3; mult_sel = ap.decay_mult; }
                                else { nstate = 4; mult_sel = ap.hang_decay_mult; }
                                break;
                            case 2:
                                if (ar.hang_counter == 0) { nstate = 4; mult_sel = ap.hang_decay_mult; } else upd = false;
                                break;
                            case 3: mult_sel = ap.decay_mult; break;
                            default: mult_sel = ap.hang_decay_mult; break;
                            }
                        }
                    }
                    ar.state = nstate;
                    if (upd) ar.volts = EX ? __fadd_rn(ar.volts, __fmul_rn(dv, mult_sel)) : fmaf(dv, mult_sel, ar.volts);
                    if (ar.volts < ap.min_volts) { ar.volts = ap.min_volts; ar.action = 0; } else { ar.action = 1; }
                    if (EX) {
                        float vo = log10f_fast(__fmul_rn(ap.inv_max_input, ar.volts));
                        if (vo > 0.0f) vo = 0.0f;
                        ad[i] = __fmul_rn(out_sample, __fdiv_rn(__fsub_rn(ap.out_target, __fmul_rn(ap.slope_constant, vo)), ar.volts));
                    } else {
                        // gain law (:563-570): Math_log10f_fast of inv_max_input * volts by bit operations, fast division
                        const unsigned ub = __float_as_uint(__fmul_rn(ap.inv_max_input, ar.volts));
                        const float F = __uint_as_float((ub & 0x007fffffu) | 0x3f000000u);
                        const float E = (float)((int)(ub >> 23) - 126);
                        float Y = fmaf(1.23149591368684f, F, -4.11852516267426f);
                        Y = fmaf(Y, F, 6.02197014179219f);
                        Y = fmaf(Y, F, -3.13396450166353f);
                        float vo = __fmul_rn(__fadd_rn(Y, E), 0.3010299956639812f);
                        vo = fminf(vo, 0.0f);
                        ad[i] = __fmul_rn(out_sample, __fdividef(fmaf(-ap.slope_constant, vo, ap.out_target), ar.volts));
                    }
                    sm[i] = fabsf(x);
                }
                ar.hang_action = (ar.hang_backaverage > ap.hang_level) ? 1 : 0;
                // this block becomes block -1: its suffix maxima replace those of block -6
                float *nsf = suf + (head * ND) * S2_THREADS;
#pragma unroll
                for (int i = ND - 1; i >= 0; i--) { nsuf = fmaxf(nsuf, sm[i]); nsf[i * S2_THREADS] = nsuf; }
                head = (head + 1 == NSUF) ? 0 : head + 1;
                m5 = 0.0f;
#pragma unroll
                for (int b = 1; b < NSUF; b++) { int sl = head + b; if (sl >= NSUF) sl -= NSUF; m5 = fmaxf(m5, suf[(sl * ND) * S2_THREADS]); }
                if (remove_dc) {                                  // audio_agc.c:577-594, double expression
#pragma unroll
                    for (int i = 0; i < ND; i++) {
                        const float wv = (float)((double)ad[i] + (double)agc_wold * 0.9999);
                        ad[i] = __fsub_rn(wv, agc_wold);
                        agc_wold = wv;
                    }
                }
            }
            if (phase == 1) {
                float4 *d0 = reinterpret_cast<float4 *>(sc + (size_t)blk * ND);
#pragma unroll
                for (int q = 0; q < ND / 4; q++) d0[q] = make_float4(ad[4 * q], ad[4 * q + 1], ad[4 * q + 2], ad[4 * q + 3]);
                continue;
            }
        }
        // fixed gain (:2513-2524), biquad_1 (:2527)
#pragma unroll
        for (int i = 0; i < ND; i++) {
            float x = __fmul_rn(ad[i], scale_gain);
            if constexpr (LEAN) {
                if (i == ND - 2) xl2 = x;
                if (i == ND - 1) xl1 = x;
                const float w = fmaf(bc[2][2], bs[2].x1, __fmul_rn(bc[2][4], bs[2].y1));
                const float y = fmaf(bc[2][0], x, tq[2]);
                tq[2] = fmaf(bc[2][3], y, fmaf(bc[2][1], x, w));
                bs[2].x2 = bs[2].x1; bs[2].x1 = x; bs[2].y2 = bs[2].y1; bs[2].y1 = y;
                ip[INTERP_HIST + i] = y;
                continue;
            }
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if (!ident[s]) {
                    const float w = fmaf(bc[s][2], bs[s].x1, __fmul_rn(bc[s][4], bs[s].y1));
                    const float y = fmaf(bc[s][0], x, tq[s]);
                    tq[s] = fmaf(bc[s][3], y, fmaf(bc[s][1], x, w));
                    bs[s].x2 = bs[s].x1; bs[s].x1 = x; bs[s].y2 = bs[s].y1; bs[s].y1 = y;
                    x = y;
                } else {
                    bs[s].x2 = bs[s].x1; bs[s].x1 = x; bs[s].y2 = bs[s].y1; bs[s].y1 = x;
                }
            }
            ip[INTERP_HIST + i] = x;
        }
        float o48[BLK];
        if (L == 4 && P == 4) interp_block2<L, 4, ND>(ip, ic, o48);
        else if (L == 4 && P == 1) interp_block2<L, 1, ND>(ip, ic, o48);
        else if (L == 2 && P == 8) interp_block2<L, 8, ND>(ip, ic, o48);
        else interp_block2<L, 2, ND>(ip, ic, o48);
#pragma unroll
        for (int i = 0; i < INTERP_HIST; i++) ip[i] = ip[ND + i];
        if constexpr (LEAN) {
            const bool muted = mute && mute[blk];
            int2 *dst = audio + (size_t)blk * BLK;
            ol1 = o48[BLK - 1]; ol2 = o48[BLK - 2];
#pragma unroll
            for (int n = 0; n < BLK; n += 4) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; e++) v[e] = muted ? 0.0f : __fmul_rn(o48[n + e], 10.0f);            // LINE_OUT_SCALING_FACTOR (:2860)
                const int w0 = format_audio_word(v[0]), w1 = format_audio_word(v[1]), w2 = format_audio_word(v[2]), w3 = format_audio_word(v[3]);
                asm volatile("st.global.v8.b32 [%0], {%1, %1, %2, %2, %3, %3, %4, %4};" ::"l"(dst + n), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
                if (audio_f) *reinterpret_cast<float4 *>(audio_f + (size_t)blk * BLK + n) = make_float4(v[0], v[1], v[2], v[3]);
            }
        } else {
            out_block(os, o48, mute && mute[blk], audio + (size_t)blk * BLK, audio_f ? audio_f + (size_t)blk * BLK : nullptr);
        }
    }
    };
    if (lean) run(std::true_type{}); else run(std::false_type{});
    if (lean && phase != 1 && a.nblocks > 0) {
        // skipped (pass-through) biquad stages saw the nearest computed stage's output before them; the identity treble stage its input
        bs[0].x1 = xl1; bs[0].x2 = xl2; bs[0].y1 = xl1; bs[0].y2 = xl2;
        bs[1] = bs[0];
        bs[3].x1 = bs[2].y1; bs[3].x2 = bs[2].y2; bs[3].y1 = bs[2].y1; bs[3].y2 = bs[2].y2;
        os.ts.x1 = ol1; os.ts.x2 = ol2; os.ts.y1 = ol1; os.ts.y2 = ol2;
    }

    // ---- state ----
    if (phase != 2) {
        if (amsam) demod2_store(st, ds, apc);
#pragma unroll
        for (int j = 0; j < 10; j++) if (j >= ppad) st.pre_s[j - ppad] = ps[j];
        if (!agc_off) {
            for (int i = 0; i < AGC_RB; i++) st.agc_ring[i] = ring[i * S2_THREADS];
            st.agc_out_index = ar.out_index; st.agc_in_index = ar.in_index; st.agc_ring_max = ar.ring_max;
            st.agc_volts = ar.volts; st.agc_save_volts = ar.save_volts; st.agc_fast_backaverage = ar.fast_backaverage;
            st.agc_hang_backaverage = ar.hang_backaverage; st.agc_hang_counter = ar.hang_counter;
            st.agc_decay_type = ar.decay_type; st.agc_state = ar.state; st.agc_action = ar.action; st.agc_hang_action = ar.hang_action;
            st.agc_wold = agc_wold;
        }
    }
    if (phase != 1) {
#pragma unroll
        for (int i = 0; i < INTERP_HIST; i++) st.interp_hist[i] = ip[i];
#pragma unroll
        for (int s = 0; s < 4; s++) st.bq1[s] = bs[s];
        out_save(os, st);
    }
}

// ---- FM: discriminator, de-emphasis, squelch at 48 ksps; rescale; treble; formatting ----
// AudioDriver_DemodFM (audio_driver.c:1544-1737), the arithmetic of demod_fm (demod_device.cuh) operation for operation, but with
// every piece of state -- previous I/Q, de-emphasis and high-pass memories, the 6-stage squelch lattice with its coefficients, the
// three Goertzel detectors -- in registers: one sample costs its dependent chain, not a dozen local-memory round trips.
__device__ void serial2_fm(const RxArgs &a, int slot, int ch, float *smem)
{
    // block staging in shared memory, [sample][lane]: the per-sample loop stays rolled (a few hundred instructions), its inputs and
    // outputs are indexed at run time, and register arrays indexed at run time would live in local memory
    float *s_i = smem + threadIdx.x, *s_q = s_i + BLK * S2_THREADS, *s_a = s_q + BLK * S2_THREADS, *s_g = s_a + BLK * S2_THREADS;
    const ChanParams &p = a.params[ch];
    ChanState &st = a.state[ch];
    const float *__restrict__ pool = a.pool;
    float *sc = a.scratch + (size_t)slot * (size_t)a.scratch_stride;
    const size_t half = (size_t)a.nblocks * BLK;
    const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
    int2 *__restrict__ audio = reinterpret_cast<int2 *>(a.audio) + chan_base;
    float *__restrict__ audio_f = a.audio_f ? a.audio_f + chan_base : nullptr;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;
    OutStage os;
    out_init(os, p, st, pool, true);
    const float fm_scaling = p.fm_scaling;
    float i_prev = st.fm_i_prev, q_prev = st.fm_q_prev, lpf_prev = st.fm_lpf_prev, hpf_a = st.fm_hpf_prev_a, hpf_b = st.fm_hpf_prev_b, sql_avg = st.fm_sql_avg;
    int count = st.fm_count, squelched = st.fm_squelched, gcount = st.fm_gcount, tdet = st.fm_tdet, tone_detected = st.fm_tone_detected;
    float subdet = st.fm_subdet, gz[9], sk[6], sv[7], ss[6];
#pragma unroll
    for (int i = 0; i < 9; i++) gz[i] = st.fm_gz[i];
#pragma unroll
    for (int j = 0; j < 6; j++) { sk[j] = __ldg(pool + p.sql.k_off + j); sv[j] = __ldg(pool + p.sql.v_off + j); ss[j] = st.sql_s[j]; }
    sv[6] = __ldg(pool + p.sql.v_off + 6);
    const bool translate_on = p.fm_translate_on != 0, tone_det = p.fm_tone_det != 0;
    const int thr = p.fm_sql_threshold;
    float nxi[BLK], nxq[BLK];
    auto ld4 = [](float *d, const float *ptr) {
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(ptr));
    };
    auto fetch = [&](int blk) {
#pragma unroll
        for (int q = 0; q < BLK / 4; q++) { ld4(nxi + 4 * q, sc + (size_t)blk * BLK + 4 * q); ld4(nxq + 4 * q, sc + half + (size_t)blk * BLK + 4 * q); }
    };
    fetch(0);
    for (int blk = 0; blk < a.nblocks; blk++) {
        float ad[BLK];
#pragma unroll
        for (int n = 0; n < BLK; n++) { s_i[n * S2_THREADS] = nxi[n]; s_q[n * S2_THREADS] = nxq[n]; s_a[n * S2_THREADS] = 0.0f; }
        if (blk + 1 < a.nblocks) fetch(blk + 1);
        if (translate_on) {
            float first_hp = 0.0f;
            // the discriminator itself is feed-forward: four samples side by side, ahead of the recurrences
#pragma unroll 4
            for (int n = 0; n < BLK; n++) {
                const float bin = s_i[n * S2_THREADS], bqn = s_q[n * S2_THREADS];
                const float ip = n ? s_i[(n - 1) * S2_THREADS] : i_prev, qp = n ? s_q[(n - 1) * S2_THREADS] : q_prev;
                const float y = __fsub_rn(__fmul_rn(ip, bqn), __fmul_rn(bin, qp));
                const float x = __fadd_rn(__fmul_rn(ip, bin), __fmul_rn(bqn, qp));
                s_g[n * S2_THREADS] = atan2f(y, x);
            }
            i_prev = s_i[(BLK - 1) * S2_THREADS]; q_prev = s_q[(BLK - 1) * S2_THREADS];
#pragma unroll 1
            for (int n = 0; n < BLK; n++) {
                const float angle = s_g[n * S2_THREADS];
                // squelch noise high-pass (6-stage lattice, arm_iir_lattice_f32) on the raw angle (:1594)
                float f = angle, acc = 0.0f, fn = 0.0f;
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const float g = ss[j];
                    fn = __fsub_rn(f, __fmul_rn(sk[j], g));
                    const float gn = __fadd_rn(__fmul_rn(fn, sk[j]), g);
                    acc = __fadd_rn(acc, __fmul_rn(gn, sv[j]));
                    if (j > 0) ss[j - 1] = gn;
                    f = fn;
                }
                acc = __fadd_rn(acc, __fmul_rn(fn, sv[6]));
                ss[5] = fn;
                if (n == 0) first_hp = acc;
                const float av = (float)((double)lpf_prev + (0.05 * (double)__fsub_rn(angle, lpf_prev)));      // de-emphasis (:1566)
                lpf_prev = av;
                if (tone_det) {
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const float b0 = __fadd_rn(__fsub_rn(__fmul_rn(p.fm_gz_r[k], gz[3 * k + 1]), gz[3 * k + 2]), av);
                        gz[3 * k + 2] = gz[3 * k + 1]; gz[3 * k + 1] = b0; gz[3 * k] = b0;
                    }
                }
                if ((!squelched && !tone_det) || (tone_detected && tone_det) || !thr) {                       // audio gate :1571-1587
                    const float hb = (float)(0.96 * (double)__fsub_rn(__fadd_rn(hpf_b, av), hpf_a));
                    hpf_a = av; hpf_b = hb;
                    s_a[n * S2_THREADS] = hb;
                }
            }
            sql_avg = (float)(((double)(1 - 0.005) * (double)sql_avg) + (0.005 * (double)__fsqrt_rn(fabsf(first_hp))));
            count = (count + 1) % 200;
            if (count == 0) {
                if ((double)sql_avg > 0.175) sql_avg = (float)0.175;
                float s_ = __fmul_rn(sql_avg, 172.0f);
                if (s_ > 24.0f) s_ = 24.0f;
                s_ = __fsub_rn(22.0f, s_);
                if (thr == 0) squelched = 0;
                else if (squelched) { if (s_ >= (float)(thr + 3)) squelched = 0; }
                else if (thr > 3) { if (s_ < (float)(thr - 3)) squelched = 1; }
                else { if (s_ < (float)thr) squelched = 1; }
            }
            if (tone_det) {
                gcount++;
                if (gcount >= 400) {
                    float en[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const float ea = __fsub_rn(gz[3 * k + 1], __fmul_rn(gz[3 * k + 2], p.fm_gz_cos[k]));
                        const float eb = __fmul_rn(gz[3 * k + 2], p.fm_gz_sin[k]);
                        gz[3 * k] = 0.0f; gz[3 * k + 1] = 0.0f; gz[3 * k + 2] = 0.0f;
                        en[k] = __fsqrt_rn(__fadd_rn(__fmul_rn(ea, ea), __fmul_rn(eb, eb)));
                    }
                    const float s_ = __fadd_rn(en[0], en[1]), r_ = en[2];
                    subdet = (float)(((1 - 0.9) * (double)subdet) + ((double)__fdiv_rn(r_, __fdiv_rn(s_, 2.0f)) * 0.9));
                    if ((double)subdet > 1.75) { tdet++; if (tdet > 5) tdet = 5; }
                    else if (tdet) tdet--;
                    tone_detected = tdet >= 2 ? 1 : 0;
                    gcount = 0;
                }
            } else {
                tone_detected = 1;
            }
        }
        const bool signal_active = !squelched;
#pragma unroll
        for (int n = 0; n < BLK; n++) ad[n] = __fmul_rn(s_a[n * S2_THREADS], fm_scaling);      // rescale only (:2819-2828)
        out_block(os, ad, (mute && mute[blk]) || !signal_active, audio + (size_t)blk * BLK, audio_f ? audio_f + (size_t)blk * BLK : nullptr);
    }
    st.fm_i_prev = i_prev; st.fm_q_prev = q_prev; st.fm_lpf_prev = lpf_prev; st.fm_hpf_prev_a = hpf_a; st.fm_hpf_prev_b = hpf_b; st.fm_sql_avg = sql_avg;
    st.fm_count = count; st.fm_squelched = squelched; st.fm_gcount = gcount; st.fm_tdet = tdet; st.fm_tone_detected = tone_detected; st.fm_subdet = subdet;
#pragma unroll
    for (int i = 0; i < 9; i++) st.fm_gz[i] = gz[i];
#pragma unroll
    for (int j = 0; j < 6; j++) st.sql_s[j] = ss[j];
    out_save(os, st);
}

}  // namespace

__global__ void __launch_bounds__(S2_THREADS)
rx_serial2_kernel(RxArgs a, int phase)
{
    extern __shared__ __align__(16) float s2_smem[];
    const int slot = blockIdx.x * S2_THREADS + threadIdx.x;
    if (slot >= a.num_items) return;
    const int ch = a.chan_list[slot];
    const ChanParams &p = a.params[ch];
    const bool ex = p.topo == TOPO_AM_SAM;
    if (p.topo == TOPO_FM) { if (phase != 2) serial2_fm(a, slot, ch, s2_smem); }      // FM has no AGC hand-off: the whole chain is "phase 1"
    else if (p.M == 4) { if (ex) serial2_dec<8, true>(a, phase, slot, ch, s2_smem); else serial2_dec<8, false>(a, phase, slot, ch, s2_smem); }
    else { if (ex) serial2_dec<16, true>(a, phase, slot, ch, s2_smem); else serial2_dec<16, false>(a, phase, slot, ch, s2_smem); }
}

// chains this kernel's AGC decomposition and interpolator variants cover; anything else stays on rx_serial_kernel
bool rx_serial2_eligible(const ChanParams &p)
{
    if (p.topo == TOPO_FM) return p.sql.n == 6;         // the squelch lattice is unrolled (IIR_15k_hpf, audio_driver.c:481-483)
    if (p.M != 4 && p.M != 2) return false;
    const int nd = BLK / p.M, L = p.interp_L, P = p.interp_plen;
    if (L != p.M) return false;
    if (!((L == 4 && (P == 4 || P == 1)) || (L == 2 && (P == 8 || P == 2)))) return false;
    if (p.pre.n > 10 || (p.aa.n != 0 && p.aa.n != 6)) return false;
    if (p.agc.mode != 5 && p.agc.attack_buffsize != NSUF * nd + 1) return false;
    return true;
}

cudaError_t launch_rx_serial2(const RxArgs &a, int phase, cudaStream_t stream)
{
    if (a.num_items <= 0) return cudaSuccess;
    if (a.scratch == nullptr || a.chan_list == nullptr) return cudaErrorInvalidValue;
    // AGC ring + suffix maxima + SAM all-pass chains; phase 2 (behind the AGC) uses none of them
    const size_t smem = phase == 2 ? 0 : (size_t)((AGC_RB + NSUF * 16) * S2_THREADS + AP_FLOATS) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(rx_serial2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (a.num_items + S2_THREADS - 1) / S2_THREADS;
    rx_serial2_kernel<<<grid, S2_THREADS, smem, stream>>>(a, phase);
    return cudaGetLastError();
}

}  // namespace uhsdr

#else   // UHSDR_EXACT: the reference-order kernel of rx_serial.cu is the only serial kernel

namespace uhsdr {
bool rx_serial2_eligible(const ChanParams &) { return false; }
cudaError_t launch_rx_serial2(const RxArgs &, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace uhsdr

#endif
