// rx_generic.cu -- the general receiver kernel: one warp owns one channel for the whole launch.
//
// It runs every signal-flow topology of AudioDriver_RxProcessor
// (mchf-eclipse/drivers/audio/audio_driver.c:2603-2942): narrow SSB/CW (decimate -> Hilbert pair),
// wide SSB (Hilbert pair -> decimate), AM / synchronous AM, FM, with or without the spectral noise
// reduction, for any filter path.  Time-parallel stages (sample formatting, IQ correction,
// frequency translation, FIR filters, interpolation, output formatting) are spread over the 32
// lanes; the sample-serial recurrences (lattice IIR, AGC, biquads, PLL, discriminator) run on
// lane 0 with their state in shared memory.  The tuned kernel for the benchmark configuration
// (narrow SSB, rx_ssb_fused.cu) covers the common case faster; this one is the complete one.
#include "dsp_device.cuh"
#include "kernels.h"
#include "nr_device.cuh"

namespace uhsdr {

static constexpr int G_WARPS = 4;

struct WarpWork {
    float xi[H1 + CHUNK], xq[H1 + CHUNK];   // stage-1 input @48k: [history | new]
    float bi[H2 + CHUNK], bq[H2 + CHUNK];   // stage-2 input:      [history | new]
    float aud[CHUNK];                       // demodulated audio at the decimated rate
    float ip[INTERP_HIST + CHUNK];          // interpolator input [history | new]
    float out[CHUNK];                       // 48 ksps audio
    float scr[2 * BLK];                     // scratch (NCO values)
    float fft[512];                         // spectral-NR FFT frame
    ChanState st;
};

// y[m] = sum_k c[k] * x[base + m*M + k], k ascending (arm_fir_f32.c:522-529,
// arm_fir_decimate_f32.c:470-490); lanes take outputs m = lane, lane+32, ...
__device__ __forceinline__ float fir_dot(const float *x, const float *__restrict__ c, int ntaps)
{
    float acc = 0.0f;
    for (int k = 0; k < ntaps; k++) acc = mad(x[k], __ldg(c + k), acc);
    return acc;
}

// Reference-order arithmetic in both builds.  Used for AM / synchronous AM: their post-AGC DC
// remover (audio_agc.c:577-594, pole 0.9999) holds a state ~1e4 x the carrier level, so the
// rounding of that state -- and with it the output at the 1e-4 level -- follows every last bit of
// its input; only bit-identical inputs reproduce the reference's output within tolerance.
__device__ __forceinline__ float fir_dot_exact(const float *x, const float *__restrict__ c, int ntaps)
{
    float acc = 0.0f;
    for (int k = 0; k < ntaps; k++) acc = __fadd_rn(acc, __fmul_rn(x[k], __ldg(c + k)));
    return acc;
}

__device__ __forceinline__ void shift_history(float *buf, int H, int nnew, int lane)
{
    // keep the newest H samples: buf[0..H) = buf[nnew..nnew+H)
    float tmp[(H2 + 31) / 32];
    int cnt = 0;
    for (int i = lane; i < H; i += 32) tmp[cnt++] = buf[nnew + i];
    __syncwarp();
    cnt = 0;
    for (int i = lane; i < H; i += 32) buf[i] = tmp[cnt++];
    __syncwarp();
}

// FRONT = false: the whole chain.  FRONT = true (split path): front end and FIR stages only; what the
// sample-serial stages need goes to a.scratch, [slot][scratch_stride] floats per launch:
//   SSB topologies   demodulated audio at the decimated rate           [nblocks * 32 / M]
//   AM / SAM         decimated I, then decimated Q                     2 x [nblocks * 32 / M]
//   FM               low-passed I, then Q at 48 ksps                   2 x [nblocks * 32]
template <bool FRONT>
__global__ void __launch_bounds__(32 * G_WARPS)
rx_generic_kernel(RxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * G_WARPS + warp;
    if (slot >= a.num_items) return;
    const int ch = a.chan_list ? a.chan_list[slot] : slot;
    WarpWork &w = reinterpret_cast<WarpWork *>(smem_raw)[warp];
    const ChanParams &p = a.params[ch];
    ChanState *gst = a.state + ch;
    const float *__restrict__ pool = a.pool;

    // ---- load state -----------------------------------------------------------------------
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(gst);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&w.st);
        for (int i = lane; i < (int)(sizeof(ChanState) / 4); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    for (int i = lane; i < H1; i += 32) { w.xi[i] = w.st.s1_hist_i[i]; w.xq[i] = w.st.s1_hist_q[i]; }
    for (int i = lane; i < H2; i += 32) { w.bi[i] = w.st.s2_hist_i[i]; w.bq[i] = w.st.s2_hist_q[i]; }
    if (lane < INTERP_HIST) w.ip[lane] = w.st.interp_hist[lane];
    __syncwarp();

    ChanState &st = w.st;
    const int M = p.M;
    const int ndec_blk = BLK / M;
    NrState *nr = (p.nr_enable && a.nr) ? (a.nr + ch) : nullptr;
    float *spec_ring = (p.spectrum_enable && a.spec_ring) ? (a.spec_ring + (size_t)ch * 1024) : nullptr;

    const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
    const int2 *__restrict__ iq = reinterpret_cast<const int2 *>(a.iq) + chan_base;
    int2 *__restrict__ audio = reinterpret_cast<int2 *>(a.audio) + chan_base;
    float *__restrict__ audio_f = a.audio_f ? a.audio_f + chan_base : nullptr;
    const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;

    int clip_q = 0, clip_h = 0, clip_f = 0;
    if (lane == 0 && p.shift_kind != 0 && st.conversion_freq != p.shift_freq) {
        // FreqShift re-prepares the NCO whenever the shift frequency changes (freq_shift.c:289-305)
        st.conversion_freq = p.shift_freq; st.osc_vect_i = 0.0f; st.osc_vect_q = 1.0f;
    }
    __syncwarp();

    for (int blk0 = 0; blk0 < a.nblocks; blk0 += CHUNK_BLOCKS) {
        const int nb = min(CHUNK_BLOCKS, a.nblocks - blk0);
        const int ns = nb * BLK;
        const int ndec = nb * ndec_blk;

        // ---- front end: format, IQ correction, spectrum tap, frequency translation ----------
        for (int b = 0; b < nb; b++) {
            const int2 s = iq[(size_t)(blk0 + b) * BLK + lane];
            // audio_driver.c:2660-2685
            const int level = abs(s.x) >> 16;
            clip_q |= (level > 4096 / 4); clip_h |= (level > 4096 / 2); clip_f |= (level > 4096);
            float fi = __fmul_rn((float)s.x, 0.0000152587890625f);
            float fq = __fmul_rn((float)s.y, 0.0000152587890625f);

            if (p.iq_auto) {
                // audio_driver.c:2274-2313 (Moseley & Slump): block statistics, EMA in double
                float t1 = __fmul_rn(sign_new(fi), fq), t2 = __fmul_rn(sign_new(fi), fi), t3 = __fmul_rn(sign_new(fq), fq);
                float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
                if (UHSDR_EXACT || p.topo == TOPO_AM_SAM) {
                    for (int j = 0; j < 32; j++) {
                        s1 = __fadd_rn(s1, __shfl_sync(0xffffffffu, t1, j));
                        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, t2, j));
                        s3 = __fadd_rn(s3, __shfl_sync(0xffffffffu, t3, j));
                    }
                } else {
                    s1 = t1; s2 = t2; s3 = t3;
                    for (int d = 16; d > 0; d >>= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
                        s3 += __shfl_xor_sync(0xffffffffu, s3, d);
                    }
                }
                const float te1 = (float)(-0.003 * (double)__fdiv_rn(s1, 32.0f) + 0.997 * (double)st.teta1_old);
                const float te2 = (float)(0.003 * (double)__fdiv_rn(s2, 32.0f) + 0.997 * (double)st.teta2_old);
                const float te3 = (float)(0.003 * (double)__fdiv_rn(s3, 32.0f) + 0.997 * (double)st.teta3_old);
                const float c1 = (te2 != 0.0f) ? __fdiv_rn(te1, te2) : 0.0f;
                float help = __fmul_rn(te2, te2);
                if (help > 0.0f) help = __fdiv_rn(__fsub_rn(__fmul_rn(te3, te3), __fmul_rn(te1, te1)), help);
                const float c2 = (help > 0.0f) ? __fsqrt_rn(help) : 1.0f;
                __syncwarp();
                if (lane == 0) { st.teta1_old = te1; st.teta2_old = te2; st.teta3_old = te3; st.M_c1 = c1; st.M_c2 = c2; }
                __syncwarp();
                fq = __fadd_rn(fq, __fmul_rn(c1, fi));
                fi = __fmul_rn(fi, c2);
            } else {
                // manual gain + phase, audio_driver.c:2259-2267, :1776-1801
                fi = __fmul_rn(fi, p.adj_i);
                fq = __fmul_rn(fq, p.adj_q);
                if (p.phase_bal < 0.0f) fq = __fadd_rn(fq, __fmul_rn(fi, p.phase_bal));
                else if (p.phase_bal > 0.0f) fi = __fadd_rn(fi, __fmul_rn(fq, p.phase_bal));
            }
            // AudioDriver_SpectrumNoZoomProcessSamples, audio_driver.c:1811-1849
            if (spec_ring && p.zoom_m == 0) {
                // samp_ptr is always even and < 1022+2: 32 pairs per block, wrap when ptr >= 1023
                uint32_t ptr = st.samp_ptr + 2u * (uint32_t)lane;
                if (ptr >= 1024u) ptr -= 1024u;
                spec_ring[ptr] = fq; spec_ring[ptr + 1] = fi;
                __syncwarp();
                if (lane == 0) { uint32_t np = st.samp_ptr + 64u; if (np >= 1024u) np -= 1024u; st.samp_ptr = np; }
                __syncwarp();
            }
            // FreqShift, freq_shift.c:275-331
            if (p.shift_kind == 1) {
                // FreqShift_QuarterFs :219-262 (roles of I and Q swap for shift > 0)
                float ib = p.shift_down ? fq : fi, qb = p.shift_down ? fi : fq;
                const int ph = lane & 3;
                float ni = ib, nq = qb;
                if (ph == 1) { ni = qb; nq = -ib; }
                else if (ph == 2) { ni = -ib; nq = -qb; }
                else if (ph == 3) { ni = -qb; nq = ib; }
                if (p.shift_down) { fq = ni; fi = nq; } else { fi = ni; fq = nq; }
            } else if (p.shift_kind == 2) {
                // FreqShift_Approx :57-108: recursive oscillator, renormalised once per block
                if (lane == 0) {
                    float vq = st.osc_vect_q, vi = st.osc_vect_i;
                    for (int n = 0; n < BLK; n++) {
                        const float oq = __fsub_rn(__fmul_rn(vq, p.osc_cos), __fmul_rn(vi, p.osc_sin));
                        const float oi = __fadd_rn(__fmul_rn(vi, p.osc_cos), __fmul_rn(vq, p.osc_sin));
                        w.scr[n] = oq; w.scr[BLK + n] = oi;
                        vq = oq; vi = oi;
                    }
                    const float g = __fdiv_rn(__fsub_rn(3.0f, __fadd_rn(__fmul_rn(vq, vq), __fmul_rn(vi, vi))), 2.0f);
                    st.osc_vect_q = __fmul_rn(g, vq); st.osc_vect_i = __fmul_rn(g, vi);
                }
                __syncwarp();
                const float oq = w.scr[lane], oi = w.scr[BLK + lane];
                float ib = p.shift_down ? fq : fi, qb = p.shift_down ? fi : fq;
                const float nq = __fsub_rn(__fmul_rn(qb, oq), __fmul_rn(ib, oi));
                const float ni = __fadd_rn(__fmul_rn(ib, oq), __fmul_rn(qb, oi));
                if (p.shift_down) { fq = ni; fi = nq; } else { fi = ni; fq = nq; }
                __syncwarp();
            }
            // AudioDriver_SpectrumZoomProcessSamples, audio_driver.c:1860-1909: 4-stage DF1 low-pass on I and on Q (after the
            // translation), 4-tap FIR decimation by 2^magnify, 32 >> magnify (Q, I) pairs into the ring.  The biquads are
            // recurrences: lane 0 takes I, lane 1 takes Q; reference operation order (the spectrum is compared to 1e-4).
            if (spec_ring && p.zoom_m != 0) {
                __syncwarp();
                w.scr[lane] = fi; w.scr[BLK + lane] = fq;
                __syncwarp();
                const int M = 1 << p.zoom_m, nout = BLK >> p.zoom_m;
                if (lane < 2) {
                    float *buf = w.scr + lane * BLK;
                    BiquadS *bs = lane ? st.zoom_bq_q : st.zoom_bq_i;
                    const float *zc = pool + p.zoom_bq_off;
                    for (int sg = 0; sg < 4; sg++) {
                        const float b0 = __ldg(zc + 5 * sg), b1 = __ldg(zc + 5 * sg + 1), b2 = __ldg(zc + 5 * sg + 2), a1 = __ldg(zc + 5 * sg + 3), a2 = __ldg(zc + 5 * sg + 4);
                        BiquadS s = bs[sg];
                        for (int i = 0; i < BLK; i++) {
                            const float x = buf[i];
                            const float acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(b0, x), __fmul_rn(b1, s.x1)), __fmul_rn(b2, s.x2)), __fmul_rn(a1, s.y1)), __fmul_rn(a2, s.y2));
                            s.x2 = s.x1; s.x1 = x; s.y2 = s.y1; s.y1 = acc;
                            buf[i] = acc;
                        }
                        bs[sg] = s;
                    }
                    // arm_fir_decimate_f32: output j = sum_k c[k] S[j M + k] over S = 3 old samples ++ the block
                    float *hist = lane ? st.zoom_hist_q : st.zoom_hist_i;
                    const float *dc = pool + p.zoom_dec_off;
                    const float c0 = __ldg(dc), c1 = __ldg(dc + 1), c2 = __ldg(dc + 2), c3 = __ldg(dc + 3);
                    const float h0 = hist[0], h1 = hist[1], h2 = hist[2];
                    hist[0] = buf[BLK - 3]; hist[1] = buf[BLK - 2]; hist[2] = buf[BLK - 1];
                    float outv[BLK / 2];
                    for (int j = 0; j < nout; j++) {
                        const int o = j * M;
                        const float s0 = o >= 3 ? buf[o - 3] : (o == 0 ? h0 : (o == 1 ? h1 : h2));
                        const float s1 = o >= 2 ? buf[o - 2] : (o == 0 ? h1 : h2);
                        const float s2 = o >= 1 ? buf[o - 1] : h2;
                        const float s3 = buf[o];
                        const float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s0, c0), __fmul_rn(s1, c1)), __fmul_rn(s2, c2)), __fmul_rn(s3, c3));
                        outv[j] = y;
                    }
                    for (int j = 0; j < nout; j++) buf[j] = outv[j];
                }
                __syncwarp();
                if (lane < nout) {
                    uint32_t ptr = st.samp_ptr + 2u * (uint32_t)lane;
                    if (ptr >= 1024u) ptr -= 1024u;
                    spec_ring[ptr] = w.scr[BLK + lane]; spec_ring[ptr + 1] = w.scr[lane];
                }
                __syncwarp();
                if (lane == 0) { uint32_t np = st.samp_ptr + 2u * (uint32_t)nout; if (np >= 1024u) np -= 1024u; st.samp_ptr = np; }
                __syncwarp();
            }
            w.xi[H1 + b * BLK + lane] = fi;
            w.xq[H1 + b * BLK + lane] = fq;
        }
        __syncwarp();

        // ---- stage 1 FIR pair @48k --------------------------------------------------------
        // output m uses inputs [m*M - (N-1), m*M] of the new samples: CMSIS appends M new samples
        // but sums from the start of its state buffer (arm_fir_decimate_f32.c:455-486)
        {
            const int N = p.s1_ntaps, M1 = p.s1_M;
            const int nout = ns / M1;
            const float *ci = pool + p.s1_ci, *cq = pool + p.s1_cq;
            const int base = H1 - (N - 1);
            if (p.topo == TOPO_AM_SAM) {
                for (int m = lane; m < nout; m += 32) {
                    w.bi[H2 + m] = fir_dot_exact(w.xi + base + m * M1, ci, N);
                    w.bq[H2 + m] = fir_dot_exact(w.xq + base + m * M1, cq, N);
                }
            } else {
                for (int m = lane; m < nout; m += 32) {
                    w.bi[H2 + m] = fir_dot(w.xi + base + m * M1, ci, N);
                    w.bq[H2 + m] = fir_dot(w.xq + base + m * M1, cq, N);
                }
            }
        }
        __syncwarp();
        shift_history(w.xi, H1, ns, lane);
        shift_history(w.xq, H1, ns, lane);

        // ---- stage 2 + demodulation ---------------------------------------------------------
        int signal_active_mask = 0xf;     // per block: FM squelch may mute
        if (p.topo == TOPO_SSB_DEC_FIRST) {
            // Hilbert pair at the decimated rate, then USB = I + Q / LSB = I - Q (:2751-2790)
            const int N = p.s2_ntaps;
            const float *ci = pool + p.s2_ci, *cq = pool + p.s2_cq;
            const int base = H2 - (N - 1);
            for (int m = lane; m < ndec; m += 32) {
                const float yi = fir_dot(w.bi + base + m, ci, N);
                const float yq = fir_dot(w.bq + base + m, cq, N);
                w.aud[m] = p.lsb ? __fsub_rn(yi, yq) : __fadd_rn(yi, yq);
            }
            __syncwarp();
            shift_history(w.bi, H2, ndec, lane);
            shift_history(w.bq, H2, ndec, lane);
        } else if (p.topo == TOPO_SSB_HIL_FIRST) {
            // combine at 48k, then decimate the audio (:2781-2803)
            for (int n = lane; n < ns; n += 32) {
                const float yi = w.bi[H2 + n], yq = w.bq[H2 + n];
                w.bi[H2 + n] = p.lsb ? __fsub_rn(yi, yq) : __fadd_rn(yi, yq);
            }
            __syncwarp();
            const int N = p.s2_ntaps, M2 = p.s2_M;
            const float *cd = pool + p.s2_ci;
            const int base = H2 - (N - 1);
            for (int m = lane; m < ndec; m += 32) w.aud[m] = fir_dot(w.bi + base + m * M2, cd, N);
            __syncwarp();
            shift_history(w.bi, H2, ns, lane);
        } else if (p.topo == TOPO_AM_SAM) {
            if constexpr (!FRONT) { if (lane == 0) demod_am_sam(p, st, pool, w.bi + H2, w.bq + H2, w.aud, nb, ndec_blk); }
            __syncwarp();
        } else {   // TOPO_FM
            if constexpr (!FRONT) {
                if (lane == 0) signal_active_mask = demod_fm(p, st, pool, w.bi + H2, w.bq + H2, w.aud, nb);
                signal_active_mask = __shfl_sync(0xffffffffu, signal_active_mask, 0);
            }
        }
        __syncwarp();
        if constexpr (FRONT) {
            float *sc = a.scratch + (size_t)slot * (size_t)a.scratch_stride;
            if (p.topo == TOPO_AM_SAM) {
                const size_t half = (size_t)a.nblocks * ndec_blk, o = (size_t)blk0 * ndec_blk;
                for (int m = lane; m < ndec; m += 32) { sc[o + m] = w.bi[H2 + m]; sc[half + o + m] = w.bq[H2 + m]; }
            } else if (p.topo == TOPO_FM) {
                const size_t half = (size_t)a.nblocks * BLK, o = (size_t)blk0 * BLK;
                for (int n = lane; n < ns; n += 32) { sc[o + n] = w.bi[H2 + n]; sc[half + o + n] = w.bq[H2 + n]; }
            } else {
                const size_t o = (size_t)blk0 * ndec_blk;
                for (int m = lane; m < ndec; m += 32) sc[o + m] = w.aud[m];
            }
            __syncwarp();
            continue;
        }

        // ---- audio post-processing, block by block (RxProcessor_DemodAudioPostprocessing) ----
        for (int b = 0; b < nb; b++) {
            float *ad = w.aud + b * ndec_blk;       // decimated block (FM: 48k block)
            float *o48 = w.out + b * BLK;
            if (p.topo != TOPO_FM) {
                if (lane == 0) {
                    // lattice pre-filter :2473-2475, AGC :2485
                    AgcRun ar = { st.agc_out_index, st.agc_in_index, st.agc_ring_max, st.agc_volts, st.agc_save_volts,
                                  st.agc_fast_backaverage, st.agc_hang_backaverage, st.agc_hang_counter, st.agc_decay_type,
                                  st.agc_state, st.agc_action, st.agc_hang_action };
                    for (int i = 0; i < ndec_blk; i++) {
                        float x = ad[i];
                        if (p.pre.n > 0) x = lattice_step(x, st.pre_s, pool + p.pre.k_off, pool + p.pre.v_off, p.pre.n);
                        if (p.agc.mode == 5) x = __fmul_rn(x, p.agc.fixed_gain);
                        else x = agc_step(x, p.agc, ar, st.agc_ring);
                        ad[i] = x;
                    }
                    if (p.agc.remove_dc && p.agc.mode != 5) {
                        // audio_agc.c:577-594: w = x + wold*0.9999 evaluated in double
                        for (int i = 0; i < ndec_blk; i++) {
                            const float wv = (float)((double)ad[i] + (double)st.agc_wold * 0.9999);
                            ad[i] = __fsub_rn(wv, st.agc_wold);
                            st.agc_wold = wv;
                        }
                    }
                    st.agc_out_index = ar.out_index; st.agc_in_index = ar.in_index; st.agc_ring_max = ar.ring_max;
                    st.agc_volts = ar.volts; st.agc_save_volts = ar.save_volts; st.agc_fast_backaverage = ar.fast_backaverage;
                    st.agc_hang_backaverage = ar.hang_backaverage; st.agc_hang_counter = ar.hang_counter;
                    st.agc_decay_type = ar.decay_type; st.agc_state = ar.state; st.agc_action = ar.action; st.agc_hang_action = ar.hang_action;
                }
                __syncwarp();
                if (nr) nr_block(p, *nr, pool, ad, ndec_blk, w.fft, lane);   // :2501-2509
                if (lane == 0) {
                    // fixed gain :2513-2524, biquad_1 :2527
                    for (int i = 0; i < ndec_blk; i++) {
                        float x = __fmul_rn(ad[i], p.scale_gain);
                        for (int s = 0; s < 4; s++) x = biquad_step(x, p.bq1[s], st.bq1[s]);
                        w.ip[INTERP_HIST + b * ndec_blk + i] = x;
                    }
                }
                __syncwarp();
                // arm_fir_interpolate_f32 :2560-2577: output n = i*L + j uses taps c[(L-1-j) + k*L]
                {
                    const int L = p.interp_L, P = p.interp_plen;
                    const float *c = pool + p.interp_c;
                    const int n = lane, i = n / L, j = n - i * L;
                    const float *x = w.ip + INTERP_HIST - (P - 1) + b * ndec_blk + i;
                    float sum = 0.0f;
                    for (int k = 0; k < P; k++) sum = mad(x[k], __ldg(c + (L - 1 - j) + k * L), sum);
                    o48[n] = sum;
                }
                __syncwarp();
            } else {
                // FM: rescale only (:2819-2828); the S-meter-only AGC on a_buffer[0] is not run
                o48[lane] = __fmul_rn(ad[lane], p.fm_scaling);
                __syncwarp();
            }
            if (lane == 0) {
                // anti-alias lattice :2581-2583, treble biquad :2832
                for (int i = 0; i < BLK; i++) {
                    float x = o48[i];
                    if (p.aa.n > 0 && p.topo != TOPO_FM) x = lattice_step(x, st.aa_s, pool + p.aa.k_off, pool + p.aa.v_off, p.aa.n);
                    o48[i] = biquad_step(x, p.bq2, st.bq2);
                }
            }
            __syncwarp();
            // output stage :2845-2941
            const bool muted = (mute && mute[blk0 + b]) || !((signal_active_mask >> b) & 1);
            float v = muted ? 0.0f : __fmul_rn(o48[lane], 10.0f);
            const int32_t word = muted ? 0 : format_audio_word(v);
            const size_t oidx = (size_t)(blk0 + b) * BLK + lane;
            audio[oidx] = make_int2(word, word);
            if (audio_f) audio_f[oidx] = v;
        }
        if (p.topo != TOPO_FM) {
            // interpolator history: keep the newest INTERP_HIST decimated samples
            __syncwarp();
            float keep = 0.0f;
            if (lane < INTERP_HIST) keep = w.ip[ndec + lane];
            __syncwarp();
            if (lane < INTERP_HIST) w.ip[lane] = keep;
        }
        __syncwarp();
    }

    // ---- store state ----------------------------------------------------------------------
    clip_q = __any_sync(0xffffffffu, clip_q); clip_h = __any_sync(0xffffffffu, clip_h); clip_f = __any_sync(0xffffffffu, clip_f);
    for (int i = lane; i < H1; i += 32) { w.st.s1_hist_i[i] = w.xi[i]; w.st.s1_hist_q[i] = w.xq[i]; }
    for (int i = lane; i < H2; i += 32) { w.st.s2_hist_i[i] = w.bi[i]; w.st.s2_hist_q[i] = w.bq[i]; }
    if (lane < INTERP_HIST) w.st.interp_hist[lane] = w.ip[lane];
    if (lane == 0) {
        st.adc_quarter_clip |= clip_q; st.adc_half_clip |= clip_h; st.adc_clip |= clip_f;
        st.blocks += a.nblocks;
    }
    __syncwarp();
    if constexpr (FRONT) {
        // only the fields this kernel owns: the serial kernels of the previous time slice may be updating the rest
        // of the record at this very moment (engine.cu runs them on a second stream)
        for (int i = lane; i < H1; i += 32) { gst->s1_hist_i[i] = w.st.s1_hist_i[i]; gst->s1_hist_q[i] = w.st.s1_hist_q[i]; }
        for (int i = lane; i < H2; i += 32) { gst->s2_hist_i[i] = w.st.s2_hist_i[i]; gst->s2_hist_q[i] = w.st.s2_hist_q[i]; }
        if (lane == 0) {
            gst->teta1_old = st.teta1_old; gst->teta2_old = st.teta2_old; gst->teta3_old = st.teta3_old;
            gst->M_c1 = st.M_c1; gst->M_c2 = st.M_c2;
            gst->osc_vect_q = st.osc_vect_q; gst->osc_vect_i = st.osc_vect_i; gst->conversion_freq = st.conversion_freq;
            gst->samp_ptr = st.samp_ptr;
            for (int k = 0; k < 4; k++) { gst->zoom_bq_i[k] = st.zoom_bq_i[k]; gst->zoom_bq_q[k] = st.zoom_bq_q[k]; gst->zoom_hist_i[k] = st.zoom_hist_i[k]; gst->zoom_hist_q[k] = st.zoom_hist_q[k]; }
            gst->adc_clip = st.adc_clip; gst->adc_half_clip = st.adc_half_clip; gst->adc_quarter_clip = st.adc_quarter_clip;
            gst->blocks = st.blocks;
        }
    } else {
        uint32_t *dst = reinterpret_cast<uint32_t *>(gst);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&w.st);
        for (int i = lane; i < (int)(sizeof(ChanState) / 4); i += 32) dst[i] = src[i];
    }
}

template <bool FRONT> static cudaError_t launch_generic(const RxArgs &a, cudaStream_t stream)
{
    // the attribute is per device: set it on every launch (an engine may live on any GPU of the process)
    const size_t smem = sizeof(WarpWork) * G_WARPS;
    cudaError_t e = cudaFuncSetAttribute(rx_generic_kernel<FRONT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (a.num_items + G_WARPS - 1) / G_WARPS;
    if (grid == 0) return cudaSuccess;
    rx_generic_kernel<FRONT><<<grid, 32 * G_WARPS, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_rx_generic(const RxArgs &a, cudaStream_t stream) { return launch_generic<false>(a, stream); }

cudaError_t launch_rx_front(const RxArgs &a, cudaStream_t stream)
{
    if (a.scratch == nullptr || a.chan_list == nullptr) return cudaErrorInvalidValue;
    return launch_generic<true>(a, stream);
}

// Spectral noise reduction of the split path (AudioDriver_RxProcessorNoiseReduction, audio_driver.c:2328-2434, with
// the deferred task of audio_nr.c run after every block): one warp per channel walks through the AGC output in
// a.scratch block by block, in place.
__global__ void __launch_bounds__(32 * G_WARPS)
rx_nr_kernel(RxArgs a)
{
    __shared__ __align__(16) float fft[G_WARPS][512];
    __shared__ float hs[G_WARPS][3 + BLK + 19 + BLK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * G_WARPS + warp;
    if (slot >= a.num_items) return;
    const int ch = a.chan_list[slot];
    const ChanParams &p = a.params[ch];
    if (!p.nr_enable || !a.nr) return;
    const int nd = BLK / p.M;
    float *sc = a.scratch + (size_t)slot * (size_t)a.scratch_stride;
    nr_slice(p, a.nr[ch], a.pool, sc, a.nblocks, nd, fft[warp], hs[warp], lane);
}

cudaError_t launch_rx_nr(const RxArgs &a, cudaStream_t stream)
{
    const int grid = (a.num_items + G_WARPS - 1) / G_WARPS;
    if (grid == 0) return cudaSuccess;
    if (a.scratch == nullptr || a.chan_list == nullptr) return cudaErrorInvalidValue;
    rx_nr_kernel<<<grid, 32 * G_WARPS, 0, stream>>>(a);
    return cudaGetLastError();
}

// scratch floats per 32-sample block: see the layout at rx_generic_kernel; 0 = the chain has a shape the serial
// kernel keeps no registers for (it stays on the general kernel)
int rx_split_floats_per_block(const ChanParams &p)
{
    if (p.pre.n > 10 || (p.aa.n != 0 && p.aa.n != 6) || p.interp_plen > INTERP_HIST + 1) return 0;
    if (p.topo == TOPO_FM) return 2 * BLK;
    if (p.topo == TOPO_AM_SAM) return 2 * (BLK / p.M);
    return BLK / p.M;
}

}  // namespace uhsdr
