// rx_serial.cu -- the sample-serial half of the split general receiver path: one channel per THREAD.
//
// launch_rx_front (rx_generic.cu, one warp per channel) has already run the time-parallel stages --
// sample formatting, IQ correction, frequency translation and every FIR of the chain -- and left the
// result in a.scratch.  What remains of AudioDriver_RxProcessor (mchf-eclipse/drivers/audio/
// audio_driver.c:2603-2942) is a set of recurrences whose state advances one sample at a time:
//   AM / synchronous AM demodulation (AudioDriver_DemodSAM :1990-2166), FM discriminator + squelch
//   (AudioDriver_DemodFM :1544-1737), lattice pre-filter (:2473-2475), WDSP AGC (audio_agc.c:349-595),
//   fixed gain + 4-stage biquad (:2513-2527), polyphase interpolator (:2560-2577), anti-alias lattice
//   (:2581-2583), treble biquad (:2832), output formatting (:2845-2941).
// Those cannot be spread over lanes, but with thousands of channels per GPU they do not have to be: every
// thread owns a channel, keeps the channel's serial state in a per-thread copy (local memory is
// lane-interleaved, so all 32 lanes touching the same field is one coalesced access) and walks through the
// launch block by block.  The arithmetic is the device functions of dsp_device.cuh / demod_device.cuh in
// the reference's order, so the exact build stays bit-exact.
#include "demod_device.cuh"
#include "kernels.h"

namespace uhsdr {

namespace {

constexpr int SER_THREADS = 32;

__device__ __forceinline__ void load_ser(SerState &s, const ChanState &g, bool notch)
{
#define CP1(f) s.f = g.f;
#define CPA(f, n) for (int i_ = 0; i_ < (n); i_++) s.f[i_] = g.f[i_];
    CPA(interp_hist, INTERP_HIST) CPA(pre_s, MAX_LAT) CPA(aa_s, MAX_LAT) CPA(sql_s, MAX_LAT) CPA(bq1, 4) CP1(bq2)
    CPA(agc_ring, AGC_RB) CP1(agc_out_index) CP1(agc_in_index) CP1(agc_ring_max) CP1(agc_volts) CP1(agc_save_volts)
    CP1(agc_fast_backaverage) CP1(agc_hang_backaverage) CP1(agc_hang_counter) CP1(agc_decay_type) CP1(agc_state) CP1(agc_wold)
    CP1(agc_action) CP1(agc_hang_action)
    CP1(sam_fil_out) CP1(sam_lowpass) CP1(sam_omega2) CP1(sam_phs) CP1(sam_dsI) CP1(sam_dsQ)
    CPA(sam_a, 24) CPA(sam_b, 24) CPA(sam_c, 24) CPA(sam_d, 24) CP1(sam_count) CP1(fade_dc27) CP1(fade_dc_insert) CP1(carrier_freq_offset)
    CP1(fm_i_prev) CP1(fm_q_prev) CP1(fm_lpf_prev) CP1(fm_hpf_prev_a) CP1(fm_hpf_prev_b) CP1(fm_sql_avg) CP1(fm_count) CP1(fm_squelched)
    CPA(fm_gz, 9) CP1(fm_subdet) CP1(fm_gcount) CP1(fm_tdet) CP1(fm_tone_detected)
    if (notch) {
        CPA(notch_coef, NOTCH_TAPS) CPA(notch_x, NOTCH_TAPS) CPA(notch_delay, NOTCH_DELAY)
        CP1(notch_energy) CP1(notch_x0) CP1(notch_head) CP1(notch_inbuf) CP1(notch_outbuf)
    }
#undef CP1
#undef CPA
}

__device__ __forceinline__ void store_ser(ChanState &g, const SerState &s, bool notch)
{
#define CP1(f) g.f = s.f;
#define CPA(f, n) for (int i_ = 0; i_ < (n); i_++) g.f[i_] = s.f[i_];
    CPA(interp_hist, INTERP_HIST) CPA(pre_s, MAX_LAT) CPA(aa_s, MAX_LAT) CPA(sql_s, MAX_LAT) CPA(bq1, 4) CP1(bq2)
    CPA(agc_ring, AGC_RB) CP1(agc_out_index) CP1(agc_in_index) CP1(agc_ring_max) CP1(agc_volts) CP1(agc_save_volts)
    CP1(agc_fast_backaverage) CP1(agc_hang_backaverage) CP1(agc_hang_counter) CP1(agc_decay_type) CP1(agc_state) CP1(agc_wold)
    CP1(agc_action) CP1(agc_hang_action)
    CP1(sam_fil_out) CP1(sam_lowpass) CP1(sam_omega2) CP1(sam_phs) CP1(sam_dsI) CP1(sam_dsQ)
    CPA(sam_a, 24) CPA(sam_b, 24) CPA(sam_c, 24) CPA(sam_d, 24) CP1(sam_count) CP1(fade_dc27) CP1(fade_dc_insert) CP1(carrier_freq_offset)
    CP1(fm_i_prev) CP1(fm_q_prev) CP1(fm_lpf_prev) CP1(fm_hpf_prev_a) CP1(fm_hpf_prev_b) CP1(fm_sql_avg) CP1(fm_count) CP1(fm_squelched)
    CPA(fm_gz, 9) CP1(fm_subdet) CP1(fm_gcount) CP1(fm_tdet) CP1(fm_tone_detected)
    if (notch) {
        CPA(notch_coef, NOTCH_TAPS) CPA(notch_x, NOTCH_TAPS) CPA(notch_delay, NOTCH_DELAY)
        CP1(notch_energy) CP1(notch_x0) CP1(notch_head) CP1(notch_inbuf) CP1(notch_outbuf)
    }
#undef CP1
#undef CPA
}

}  // namespace

// arm_fir_interpolate_f32 (:2560-2577) for one block with the interpolation factor and the polyphase length known
// at compile time: output n = i*L + j uses taps c[(L-1-j) + k*L] on ip[INTERP_HIST - (P-1) + i + k], k ascending.
template <int L, int P>
__device__ __forceinline__ void interp_block(const float *ip, const float *__restrict__ ic, float *o48)
{
    float c[L][P];
#pragma unroll
    for (int j = 0; j < L; j++)
#pragma unroll
        for (int k = 0; k < P; k++) c[j][k] = __ldg(ic + (L - 1 - j) + k * L);
    float x[P];
#pragma unroll
    for (int k = 0; k < P - 1; k++) x[k + 1] = ip[INTERP_HIST - (P - 1) + k];
#pragma unroll 4
    for (int i = 0; i < BLK / L; i++) {
#pragma unroll
        for (int k = 0; k < P - 1; k++) x[k] = x[k + 1];
        x[P - 1] = ip[INTERP_HIST + i];
#pragma unroll
        for (int j = 0; j < L; j++) {
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < P; k++) sum = mad(x[k], c[j][k], sum);
            o48[i * L + j] = sum;
        }
    }
}

// arm_iir_lattice_f32 (arm_iir_lattice_f32.c:348-440) with coefficients and state in registers, NS stage slots with
// the filter's n stages at the end: front slots have k = v = 0 and pass the sample through unchanged (bit-exact)
template <int NS>
__device__ __forceinline__ float lattice_regs(float x, const float (&k)[NS], const float (&v)[NS + 1], float (&s)[NS])
{
    float f = x, acc = 0.0f, fn = x;
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const float g = s[j];
        fn = __fsub_rn(f, __fmul_rn(k[j], g));
        const float gn = __fadd_rn(__fmul_rn(fn, k[j]), g);
        acc = __fadd_rn(acc, __fmul_rn(gn, v[j]));
        if (j > 0) s[j - 1] = gn;
        f = fn;
    }
    acc = __fadd_rn(acc, __fmul_rn(fn, v[NS]));
    s[NS - 1] = fn;
    return acc;
}

// LMS automatic notch on one decimated block, in place: AudioDriver_NotchFilter (audio_driver.c:1746-1763) around
// arm_lms_norm_f32 (CMSIS arm_lms_norm_f32.c:372-443, the portable loop).  The adaptive FIR runs on the current
// audio, the reference input is the audio delayed through the 128-sample line, and the notched audio is the error.
__device__ __forceinline__ void notch_block(SerState &st, float *buf, int bs, float mu)
{
    for (int i = 0; i < bs; i++) st.notch_delay[st.notch_inbuf + i] = buf[i];
    float energy = st.notch_energy, x0 = st.notch_x0;
    int head = st.notch_head;
    for (int i = 0; i < bs; i++) {
        const float in = buf[i];
        st.notch_x[(head + NOTCH_TAPS - 1) & (NOTCH_TAPS - 1)] = in;        // window = the newest 64 inputs, oldest at head
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
    st.notch_inbuf += bs;
    st.notch_outbuf = st.notch_inbuf + bs;
    st.notch_inbuf %= NOTCH_DELAY;
    st.notch_outbuf %= NOTCH_DELAY;
}

// phase 0: the whole serial chain.  Channels with the spectral noise reduction, which sits between the AGC and the
// biquad cascade (audio_driver.c:2501-2509) and works on warp-cooperative FFT frames (rx_nr_kernel), run it in two
// launches: phase 1 = demodulation, lattice, AGC, result back into a.scratch; phase 2 = everything after the NR.
__global__ void __launch_bounds__(SER_THREADS)
rx_serial_kernel(RxArgs a, int phase)
{
    const int slot = blockIdx.x * SER_THREADS + threadIdx.x;
    if (slot >= a.num_items) return;
    const int ch = a.chan_list[slot];
    const ChanParams &p = a.params[ch];
    const float *__restrict__ pool = a.pool;
    SerState st;
    const bool notch = p.notch_enable != 0;
    load_ser(st, a.state[ch], notch);

    const int M = p.M;
    const int nd = BLK / M;                         // decimated samples per block (FM: M = 1, unused)
    const bool fm = p.topo == TOPO_FM, amsam = p.topo == TOPO_AM_SAM;
    float *sc = a.scratch + (size_t)slot * (size_t)a.scratch_stride;
    const size_t half = fm ? (size_t)a.nblocks * BLK : (size_t)a.nblocks * nd;
    const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
    int2 *__restrict__ audio = reinterpret_cast<int2 *>(a.audio) + chan_base;
    float *__restrict__ audio_f = a.audio_f ? a.audio_f + chan_base : nullptr;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;

    // ---- coefficients and the small recurrences' state in registers ----
    const AgcP ap = p.agc;
    AgcRun ar = { st.agc_out_index, st.agc_in_index, st.agc_ring_max, st.agc_volts, st.agc_save_volts,
                  st.agc_fast_backaverage, st.agc_hang_backaverage, st.agc_hang_counter, st.agc_decay_type,
                  st.agc_state, st.agc_action, st.agc_hang_action };
    float pk[10], pv[11], ps[10];                   // lattice pre-filter, front-padded to 10 stages
    const int pn = p.pre.n, ppad = 10 - pn;
#pragma unroll
    for (int j = 0; j < 10; j++) {
        pk[j] = (j >= ppad) ? __ldg(pool + p.pre.k_off + (j - ppad)) : 0.0f;
        pv[j] = (j >= ppad) ? __ldg(pool + p.pre.v_off + (j - ppad)) : 0.0f;
        ps[j] = (j >= ppad) ? st.pre_s[j - ppad] : 0.0f;
    }
    pv[10] = (pn > 0) ? __ldg(pool + p.pre.v_off + pn) : 1.0f;
    const bool aa_on = p.aa.n == 6 && !fm;
    float ak[6], av[7], as_[6];                     // anti-alias lattice (6 stages where present)
#pragma unroll
    for (int j = 0; j < 6; j++) {
        ak[j] = aa_on ? __ldg(pool + p.aa.k_off + j) : 0.0f;
        av[j] = aa_on ? __ldg(pool + p.aa.v_off + j) : 0.0f;
        as_[j] = aa_on ? st.aa_s[j] : 0.0f;
    }
    av[6] = aa_on ? __ldg(pool + p.aa.v_off + 6) : 1.0f;
    float bc[4][5], tc[5];
    BiquadS bs[4], ts = st.bq2;
#pragma unroll
    for (int s = 0; s < 4; s++) {
#pragma unroll
        for (int q = 0; q < 5; q++) bc[s][q] = p.bq1[s][q];
        bs[s] = st.bq1[s];
    }
#pragma unroll
    for (int q = 0; q < 5; q++) tc[q] = p.bq2[q];
    const float scale_gain = p.scale_gain, fm_scaling = p.fm_scaling, notch_mu = p.notch_mu;
    const bool remove_dc = ap.remove_dc && ap.mode != 5;
    float agc_wold = st.agc_wold;
    float ip[INTERP_HIST + BLK / 2];                // interpolator input: [history | decimated block]
    for (int i = 0; i < INTERP_HIST; i++) ip[i] = st.interp_hist[i];
    const int L = p.interp_L, P = p.interp_plen;
    const float *__restrict__ ic = pool + p.interp_c;

    for (int blk = 0; blk < a.nblocks; blk++) {
        float ad[BLK], o48[BLK];
        bool signal_active = true;
        // ---- demodulation ----
        if (phase == 2) {
            for (int n = 0; n < nd; n++) ad[n] = sc[(size_t)blk * nd + n];
        } else if (fm) {
            float bi[BLK], bq[BLK];
            const float4 *si = reinterpret_cast<const float4 *>(sc + (size_t)blk * BLK), *sq = reinterpret_cast<const float4 *>(sc + half + (size_t)blk * BLK);
            for (int n = 0; n < BLK / 4; n++) {
                const float4 vi = si[n], vq = sq[n];
                bi[4 * n] = vi.x; bi[4 * n + 1] = vi.y; bi[4 * n + 2] = vi.z; bi[4 * n + 3] = vi.w;
                bq[4 * n] = vq.x; bq[4 * n + 1] = vq.y; bq[4 * n + 2] = vq.z; bq[4 * n + 3] = vq.w;
            }
            signal_active = demod_fm(p, st, pool, bi, bq, ad, 1) != 0;
        } else if (amsam) {
            float bi[BLK / 2], bq[BLK / 2];
            for (int n = 0; n < nd; n++) { bi[n] = sc[(size_t)blk * nd + n]; bq[n] = sc[half + (size_t)blk * nd + n]; }
            demod_am_sam(p, st, pool, bi, bq, ad, 1, nd);
        } else {
            const float4 *sa = reinterpret_cast<const float4 *>(sc + (size_t)blk * nd);
            for (int n = 0; n < nd / 4; n++) { const float4 v = sa[n]; ad[4 * n] = v.x; ad[4 * n + 1] = v.y; ad[4 * n + 2] = v.z; ad[4 * n + 3] = v.w; }
        }
        // ---- audio post-processing (RxProcessor_DemodAudioPostprocessing) ----
        if (!fm) {
            // lattice pre-filter :2473-2475, AGC :2485 (+ DC remover audio_agc.c:577-594, double expression),
            // fixed gain :2513-2524, biquad_1 :2527
            if (phase != 2) {
                if (notch) notch_block(st, ad, nd, notch_mu);
                for (int i = 0; i < nd; i++) {
                    float x = ad[i];
                    if (pn > 0) x = lattice_regs<10>(x, pk, pv, ps);
                    if (ap.mode == 5) x = __fmul_rn(x, ap.fixed_gain);
                    else x = agc_step(x, ap, ar, st.agc_ring);
                    ad[i] = x;
                }
                if (remove_dc) {
                    for (int i = 0; i < nd; i++) {
                        const float wv = (float)((double)ad[i] + (double)agc_wold * 0.9999);
                        ad[i] = __fsub_rn(wv, agc_wold);
                        agc_wold = wv;
                    }
                }
                if (phase == 1) {
                    for (int n = 0; n < nd; n++) sc[(size_t)blk * nd + n] = ad[n];
                    continue;
                }
            }
            for (int i = 0; i < nd; i++) {
                float x = __fmul_rn(ad[i], scale_gain);
#pragma unroll
                for (int s = 0; s < 4; s++) x = biquad_step(x, bc[s], bs[s]);
                ip[INTERP_HIST + i] = x;
            }
            if (L == 4 && P == 4) interp_block<4, 4>(ip, ic, o48);
            else if (L == 4 && P == 1) interp_block<4, 1>(ip, ic, o48);
            else if (L == 2 && P == 8) interp_block<2, 8>(ip, ic, o48);
            else if (L == 2 && P == 2) interp_block<2, 2>(ip, ic, o48);
            else {
                for (int n = 0; n < BLK; n++) {
                    const int i = n / L, j = n - i * L;
                    const float *x = ip + INTERP_HIST - (P - 1) + i;
                    float sum = 0.0f;
                    for (int k = 0; k < P; k++) sum = mad(x[k], __ldg(ic + (L - 1 - j) + k * L), sum);
                    o48[n] = sum;
                }
            }
            for (int i = 0; i < INTERP_HIST; i++) ip[i] = ip[nd + i];      // keep the newest INTERP_HIST decimated samples
        } else {
            // FM: rescale only (:2819-2828); the S-meter-only AGC on a_buffer[0] is not run
            for (int n = 0; n < BLK; n++) o48[n] = __fmul_rn(ad[n], fm_scaling);
        }
        // anti-alias lattice :2581-2583, treble biquad :2832, output stage :2845-2941
        const bool muted = (mute && mute[blk]) || !signal_active;
        int2 *dst = audio + (size_t)blk * BLK;
#pragma unroll 2
        for (int n = 0; n < BLK; n += 2) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                float x = o48[n + e];
                if (aa_on) x = lattice_regs<6>(x, ak, av, as_);
                x = biquad_step(x, tc, ts);
                v[e] = muted ? 0.0f : __fmul_rn(x, 10.0f);
            }
            const int w0 = muted ? 0 : format_audio_word(v[0]), w1 = muted ? 0 : format_audio_word(v[1]);
            *reinterpret_cast<int4 *>(dst + n) = make_int4(w0, w0, w1, w1);
            if (audio_f) *reinterpret_cast<float2 *>(audio_f + (size_t)blk * BLK + n) = make_float2(v[0], v[1]);
        }
    }

    for (int i = 0; i < INTERP_HIST; i++) st.interp_hist[i] = ip[i];
#pragma unroll
    for (int j = 0; j < 10; j++) if (j >= ppad) st.pre_s[j - ppad] = ps[j];
    if (aa_on) {
#pragma unroll
        for (int j = 0; j < 6; j++) st.aa_s[j] = as_[j];
    }
#pragma unroll
    for (int s = 0; s < 4; s++) st.bq1[s] = bs[s];
    st.bq2 = ts;
    st.agc_wold = agc_wold;
    st.agc_out_index = ar.out_index; st.agc_in_index = ar.in_index; st.agc_ring_max = ar.ring_max;
    st.agc_volts = ar.volts; st.agc_save_volts = ar.save_volts; st.agc_fast_backaverage = ar.fast_backaverage;
    st.agc_hang_backaverage = ar.hang_backaverage; st.agc_hang_counter = ar.hang_counter;
    st.agc_decay_type = ar.decay_type; st.agc_state = ar.state; st.agc_action = ar.action; st.agc_hang_action = ar.hang_action;
    store_ser(a.state[ch], st, notch);
}

cudaError_t launch_rx_serial(const RxArgs &a, int phase, cudaStream_t stream)
{
    if (a.num_items <= 0) return cudaSuccess;
    if (a.scratch == nullptr || a.chan_list == nullptr) return cudaErrorInvalidValue;
    const int grid = (a.num_items + SER_THREADS - 1) / SER_THREADS;
    rx_serial_kernel<<<grid, SER_THREADS, 0, stream>>>(a, phase);
    return cudaGetLastError();
}

}  // namespace uhsdr
