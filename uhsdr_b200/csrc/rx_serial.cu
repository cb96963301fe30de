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

__device__ __forceinline__ void load_ser(SerState &s, const ChanState &g)
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
#undef CP1
#undef CPA
}

__device__ __forceinline__ void store_ser(ChanState &g, const SerState &s)
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
#undef CP1
#undef CPA
}

}  // namespace

__global__ void __launch_bounds__(SER_THREADS)
rx_serial_kernel(RxArgs a)
{
    const int slot = blockIdx.x * SER_THREADS + threadIdx.x;
    if (slot >= a.num_items) return;
    const int ch = a.chan_list[slot];
    const ChanParams &p = a.params[ch];
    const float *__restrict__ pool = a.pool;
    SerState st;
    load_ser(st, a.state[ch]);

    const int M = p.M;
    const int nd = BLK / M;                         // decimated samples per block (FM: M = 1, unused)
    const bool fm = p.topo == TOPO_FM, amsam = p.topo == TOPO_AM_SAM;
    const float *__restrict__ sc = a.scratch + (size_t)slot * (size_t)a.scratch_stride;
    const size_t half = fm ? (size_t)a.nblocks * BLK : (size_t)a.nblocks * nd;
    const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
    int2 *__restrict__ audio = reinterpret_cast<int2 *>(a.audio) + chan_base;
    float *__restrict__ audio_f = a.audio_f ? a.audio_f + chan_base : nullptr;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;

    AgcRun ar = { st.agc_out_index, st.agc_in_index, st.agc_ring_max, st.agc_volts, st.agc_save_volts,
                  st.agc_fast_backaverage, st.agc_hang_backaverage, st.agc_hang_counter, st.agc_decay_type,
                  st.agc_state, st.agc_action, st.agc_hang_action };
    float ip[INTERP_HIST + BLK];                    // interpolator input: [history | decimated block]
    for (int i = 0; i < INTERP_HIST; i++) ip[i] = st.interp_hist[i];
    const int L = p.interp_L, P = p.interp_plen;
    const float *__restrict__ ic = pool + p.interp_c;

    for (int blk = 0; blk < a.nblocks; blk++) {
        float ad[BLK], o48[BLK];
        bool signal_active = true;
        // ---- demodulation ----
        if (fm) {
            float bi[BLK], bq[BLK];
            for (int n = 0; n < BLK; n++) { bi[n] = sc[(size_t)blk * BLK + n]; bq[n] = sc[half + (size_t)blk * BLK + n]; }
            signal_active = demod_fm(p, st, pool, bi, bq, ad, 1) != 0;
        } else if (amsam) {
            float bi[BLK / 2], bq[BLK / 2];
            for (int n = 0; n < nd; n++) { bi[n] = sc[(size_t)blk * nd + n]; bq[n] = sc[half + (size_t)blk * nd + n]; }
            demod_am_sam(p, st, pool, bi, bq, ad, 1, nd);
        } else {
            for (int n = 0; n < nd; n++) ad[n] = sc[(size_t)blk * nd + n];
        }
        // ---- audio post-processing (RxProcessor_DemodAudioPostprocessing) ----
        if (!fm) {
            for (int i = 0; i < nd; i++) {
                float x = ad[i];
                if (p.pre.n > 0) x = lattice_step(x, st.pre_s, pool + p.pre.k_off, pool + p.pre.v_off, p.pre.n);
                if (p.agc.mode == 5) x = __fmul_rn(x, p.agc.fixed_gain);
                else x = agc_step(x, p.agc, ar, st.agc_ring);
                ad[i] = x;
            }
            if (p.agc.remove_dc && p.agc.mode != 5) {
                // audio_agc.c:577-594: w = x + wold*0.9999 evaluated in double
                for (int i = 0; i < nd; i++) {
                    const float wv = (float)((double)ad[i] + (double)st.agc_wold * 0.9999);
                    ad[i] = __fsub_rn(wv, st.agc_wold);
                    st.agc_wold = wv;
                }
            }
            // fixed gain :2513-2524, biquad_1 :2527
            for (int i = 0; i < nd; i++) {
                float x = __fmul_rn(ad[i], p.scale_gain);
                for (int s = 0; s < 4; s++) x = biquad_step(x, p.bq1[s], st.bq1[s]);
                ip[INTERP_HIST + i] = x;
            }
            // arm_fir_interpolate_f32 :2560-2577: output n = i*L + j uses taps c[(L-1-j) + k*L]
            for (int n = 0; n < BLK; n++) {
                const int i = n / L, j = n - i * L;
                const float *x = ip + INTERP_HIST - (P - 1) + i;
                float sum = 0.0f;
                for (int k = 0; k < P; k++) sum = mad(x[k], __ldg(ic + (L - 1 - j) + k * L), sum);
                o48[n] = sum;
            }
            for (int i = 0; i < INTERP_HIST; i++) ip[i] = ip[nd + i];      // keep the newest INTERP_HIST decimated samples
        } else {
            // FM: rescale only (:2819-2828); the S-meter-only AGC on a_buffer[0] is not run
            for (int n = 0; n < BLK; n++) o48[n] = __fmul_rn(ad[n], p.fm_scaling);
        }
        // anti-alias lattice :2581-2583, treble biquad :2832
        for (int n = 0; n < BLK; n++) {
            float x = o48[n];
            if (p.aa.n > 0 && !fm) x = lattice_step(x, st.aa_s, pool + p.aa.k_off, pool + p.aa.v_off, p.aa.n);
            o48[n] = biquad_step(x, p.bq2, st.bq2);
        }
        // output stage :2845-2941
        const bool muted = (mute && mute[blk]) || !signal_active;
        int2 *dst = audio + (size_t)blk * BLK;
        for (int n = 0; n < BLK; n += 2) {
            const float v0 = muted ? 0.0f : __fmul_rn(o48[n], 10.0f), v1 = muted ? 0.0f : __fmul_rn(o48[n + 1], 10.0f);
            const int w0 = muted ? 0 : format_audio_word(v0), w1 = muted ? 0 : format_audio_word(v1);
            *reinterpret_cast<int4 *>(dst + n) = make_int4(w0, w0, w1, w1);
            if (audio_f) *reinterpret_cast<float2 *>(audio_f + (size_t)blk * BLK + n) = make_float2(v0, v1);
        }
    }

    for (int i = 0; i < INTERP_HIST; i++) st.interp_hist[i] = ip[i];
    st.agc_out_index = ar.out_index; st.agc_in_index = ar.in_index; st.agc_ring_max = ar.ring_max;
    st.agc_volts = ar.volts; st.agc_save_volts = ar.save_volts; st.agc_fast_backaverage = ar.fast_backaverage;
    st.agc_hang_backaverage = ar.hang_backaverage; st.agc_hang_counter = ar.hang_counter;
    st.agc_decay_type = ar.decay_type; st.agc_state = ar.state; st.agc_action = ar.action; st.agc_hang_action = ar.hang_action;
    store_ser(a.state[ch], st);
}

cudaError_t launch_rx_serial(const RxArgs &a, cudaStream_t stream)
{
    if (a.num_items <= 0) return cudaSuccess;
    if (a.scratch == nullptr || a.chan_list == nullptr) return cudaErrorInvalidValue;
    const int grid = (a.num_items + SER_THREADS - 1) / SER_THREADS;
    rx_serial_kernel<<<grid, SER_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace uhsdr
