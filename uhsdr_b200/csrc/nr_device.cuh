// nr_device.cuh -- spectral noise reduction, warp-cooperative (one warp = one channel).
//   ISR side   AudioDriver_RxProcessorNoiseReduction   mchf-eclipse/drivers/audio/audio_driver.c:2328-2434
//   FIFOs      NR_in/out_buffer_*                      audio_nr.c:174-299
//   task side  AudioNr_HandleNoiseReduction            audio_nr.c:314-349 (run once after every block:
//                                                      the oracle's fixed schedule of the PendSV task)
//   algorithm  spectral_noise_reduction_3              audio_nr.c:1841-2195
#pragma once
#include "demod_device.cuh"
#include "fft_device.cuh"

namespace uhsdr {

constexpr int NR_FIFO = 5;     // NR_BUFFER_FIFO_SIZE = NR_BUFFER_NUM + 1 (freedv_uhsdr.h:130)

__device__ __forceinline__ int nr_fifo_count(int head, int tail) { int len = head - tail; return len < 0 ? len + NR_FIFO : len; }

// LPC impulse noise blanker, alt_noise_blanking (audio_nr.c:2210-2539): one 128-sample frame in place in `frame`
// (global memory), `w` = 512 floats of shared memory.  Order-10 LPC of the frame (autocorrelation + Levinson-Durbin),
// inverse filter + matched filter, threshold at nb_level/2 standard deviations x sqrt(LPC power), up to 5 impulses of
// 7 samples replaced by the windowed sum of a forward and a backward prediction; 26 samples stay in the working
// buffer between frames (output = input delayed by 13).  Every sum keeps the order of the CMSIS portable kernels
// (arm_dot_prod_f32, arm_fir_f32 with freshly zeroed state, arm_var_f32, arm_power_f32), with separate multiply and
// add in both builds: the detection is a threshold decision, so the blanker is bit-exact or it is wrong.
// Lanes: autocorrelation lag i on lane i, the two FIRs over samples, the short serial parts on lane 0.
__device__ inline void nb_frame(const ChanParams &p, NrState &nr, float *frame, float *w, int lane)
{
    constexpr int ORD = 10, PL = 3, IMP = 7;
    float *wb = w, *temp = w + 160, *temp2 = w + 288, *R = w + 416, *lpcs = w + 432, *rev = w + 448;
    int *ipos = reinterpret_cast<int *>(w + 464);              // [0] count, [1..5] positions
    for (int i = lane; i < 2 * PL + 2 * ORD; i += 32) wb[i] = nr.nb_work[i];
    for (int i = lane; i < 128; i += 32) wb[2 * PL + 2 * ORD + i] = frame[i];                 // :2344
    __syncwarp();
    const float *x = wb + ORD + PL;
    if (lane <= ORD) {                                                                        // :2356-2360
        float sum = 0.0f;
        for (int n = 0; n < 128 - lane; n++) sum = __fadd_rn(sum, __fmul_rn(x[n], x[n + lane]));
        R[lane] = sum;
    }
    __syncwarp();
    if (lane == 0) {
        float l[ORD + 1], any[ORD + 1];
        const float r0 = (float)((double)R[0] * (1.0 + 1.0e-9));                              // :2367
        l[0] = 1.0f;
        for (int i = 1; i <= ORD; i++) l[i] = 0.0f;
        float alfa = r0;
        for (int m = 1; m <= ORD; m++) {                                                      // Levinson-Durbin, :2376-2392
            float s = 0.0f;
            for (int u = 1; u < m; u++) s = __fadd_rn(s, __fmul_rn(l[u], R[m - u]));
            const float k = __fdiv_rn(-__fadd_rn(R[m], s), alfa);
            for (int v = 1; v < m; v++) any[v] = __fadd_rn(l[v], __fmul_rn(k, l[m - v]));
            for (int v = 1; v < m; v++) l[v] = any[v];
            l[m] = k;
            alfa = __fmul_rn(alfa, __fsub_rn(1.0f, __fmul_rn(k, k)));
        }
        for (int o = 0; o <= ORD; o++) { lpcs[o] = l[o]; rev[ORD - o] = l[o]; }
    }
    __syncwarp();
    for (int n = lane; n < 128; n += 32) {                                                    // inverse filter, :2402
        float acc = 0.0f;
        for (int k = 0; k <= ORD; k++) { const int j = n - ORD + k; acc = __fadd_rn(acc, __fmul_rn(j >= 0 ? x[j] : 0.0f, rev[k])); }
        temp[n] = acc;
    }
    __syncwarp();
    for (int n = lane; n < 128; n += 32) {                                                    // matched filter, :2406
        float acc = 0.0f;
        for (int k = 0; k <= ORD; k++) { const int j = n - ORD + k; acc = __fadd_rn(acc, __fmul_rn(j >= 0 ? temp[j] : 0.0f, lpcs[k])); }
        temp2[n] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        float sum = 0.0f, sumsq = 0.0f;                                                       // arm_var_f32, :2409
        for (int n = 0; n < 128; n++) { const float in = temp2[n]; sumsq = __fadd_rn(sumsq, __fmul_rn(in, in)); sum = __fadd_rn(sum, in); }
        const float sigma2 = __fdiv_rn(__fsub_rn(sumsq, __fdiv_rn(__fmul_rn(sum, sum), 128.0f)), 127.0f);
        float lpc_power = 0.0f;                                                               // arm_power_f32 over `order` values, :2411
        for (int i = 0; i < ORD; i++) lpc_power = __fadd_rn(lpc_power, __fmul_rn(lpcs[i], lpcs[i]));
        const float thr = (float)((double)(float)p.nb_level * 0.5 * (double)sqrtf(__fmul_rn(sigma2, lpc_power)));   // :2414
        int sp = ORD + PL, count = 0;
        do {                                                                                  // :2423-2437
            if ((temp2[sp] > thr) || (temp2[sp] < (-thr))) { ipos[1 + count] = sp - ORD; count++; sp += PL; }
            sp++;
        } while ((sp < 128) && (count < 5));
        ipos[0] = count;
        float nl[ORD], nrv[ORD];                                                              // negated predictors, :2446-2447
        for (int i = 0; i < ORD; i++) { nl[i] = -lpcs[1 + i]; nrv[i] = -rev[i]; }
        for (int j = 0; j < count; j++) {                                                     // :2450-2510
            const int pj = ipos[1 + j];
            float Rfw[IMP + ORD], Rbw[IMP + ORD];
            for (int k = 0; k < ORD; k++) { Rfw[k] = wb[pj + k]; Rbw[IMP + k] = wb[ORD + PL + pj + PL + k + 1]; }
            for (int i = 0; i < IMP; i++) {
                float a = 0.0f, b = 0.0f;
                for (int k = 0; k < ORD; k++) a = __fadd_rn(a, __fmul_rn(nrv[k], Rfw[i + k]));
                Rfw[i + ORD] = a;
                for (int k = 0; k < ORD; k++) b = __fadd_rn(b, __fmul_rn(nl[k], Rbw[IMP - i + k]));
                Rbw[IMP - i - 1] = b;
            }
            for (int i = 0; i < IMP; i++) {
                const float wbw = (float)(1.0 * i / (IMP - 1)), wfw = (float)(1.0 * (IMP - 1 - i) / (IMP - 1));   // :2349-2353
                wb[ORD + pj + i] = __fadd_rn(__fmul_rn(wfw, Rfw[ORD + i]), __fmul_rn(wbw, Rbw[i]));
            }
        }
    }
    __syncwarp();
    for (int i = lane; i < 128; i += 32) frame[i] = wb[ORD + PL + i];                         // :2520-2521
    for (int i = lane; i < 2 * PL + 2 * ORD; i += 32) nr.nb_work[i] = wb[128 + i];
    __syncwarp();
}

// One 128-sample frame in place in `frame` (global memory), using `fft` (512 floats of shared memory).
__device__ inline void nr_spectral(const ChanParams &p, NrState &nr, const float *__restrict__ pool, float *frame, float *fft, int lane)
{
    const float *win = pool + p.nr_win_c;
    const float psthr = 0.99f, pnsaf = 0.01f, psini = 0.5f, pspri = 0.5f;
    const float ax = 0.7405f, ap = 0.8691f;
    const float xih1 = p.nr_xih1;
    const float xih1r = (float)(1.0 / (1.0 + (double)xih1) - 1.0);
    const float pfac = (float)((1.0 / (double)pspri - 1.0) * (1.0 + (double)xih1));
    const float snr_prio_min = 0.001f;
    const float alpha = p.nr_alpha;

    if (nr.first_time == 1) {
        for (int b = lane; b < 128; b += 32) {
            nr.last_sample[b] = 0.0f; nr.Hk[b] = 1.0f; nr.Hk_old[b] = 1.0f; nr.Nest0[b] = 0.0f; nr.pslp[b] = 0.5f;
        }
        __syncwarp();
        if (lane == 0) nr.first_time = 2;
        __syncwarp();
    }
    for (int i = lane; i < 128; i += 32) {
        const float prev = nr.last_sample[i], cur = frame[i];
        fft[2 * i] = __fmul_rn(prev, __ldg(win + i)); fft[2 * i + 1] = 0.0f;
        fft[256 + 2 * i] = __fmul_rn(cur, __ldg(win + 128 + i)); fft[256 + 2 * i + 1] = 0.0f;
        nr.last_sample[i] = cur;
    }
    __syncwarp();
    fft_warp_smem<256, 8>(fft, pool + p.tw256_off, false, lane);

    // audio_nr.c:1989-2003: the 20th averaging frame switches first_time to 3 and the tracker
    // below already runs on that same frame (two consecutive ifs in the reference)
    const bool run2 = nr.first_time == 2;
    const bool run3 = nr.first_time == 3 || (run2 && nr.init_counter + 1 > 19);
    int vad_low = 0, vad_high = 63;                      // audio_nr.c:1863-1864 until first_time == 3
    if (run3) { vad_low = p.nr_vad_low; vad_high = p.nr_vad_high; }
    __syncwarp();
    for (int b = lane; b < 128; b += 32) {
        const float re = fft[2 * b], im = fft[2 * b + 1];
        const float X = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
        if (run2) {
            const float nest = (float)((double)nr.Nest0[b] + 0.05 * (double)X);
            nr.Nest0[b] = nest;
            nr.xt[b] = __fmul_rn(psini, nest);
        }
        if (run3) {
            float xt = nr.xt[b];
            // MMSE speech-presence noise tracker (:2008-2024)
            float ph1y = (float)(1.0 / (1.0 + (double)__fmul_rn(pfac, expf(__fdiv_rn(__fmul_rn(xih1r, X), xt)))));
            const float pslp = (float)((double)__fmul_rn(ap, nr.pslp[b]) + (1.0 - (double)ap) * (double)ph1y);
            nr.pslp[b] = pslp;
            if (pslp > psthr) ph1y = (float)(1.0 - (double)pnsaf);
            else ph1y = (float)fmin((double)ph1y, 1.0);
            const float xtr = (float)((1.0 - (double)ph1y) * (double)X + (double)__fmul_rn(ph1y, xt));
            xt = (float)((double)__fmul_rn(ax, xt) + (1.0 - (double)ax) * (double)xtr);
            nr.xt[b] = xt;
            // a-posteriori / a-priori SNR (:2027-2032)
            const float post = (float)fmax(fmin((double)__fdiv_rn(X, xt), 1000.0), (double)snr_prio_min);
            const float prio = (float)fmax((double)__fmul_rn(alpha, nr.Hk_old[b]) + (1.0 - (double)alpha) * fmax((double)post - 1.0, 0.0), 0.0);
            if (b >= vad_low && b < vad_high) {
                // gain (:2065-2078); musical-noise smoothing (:2080-2137) is the identity with the
                // firmware's uninitialised power_threshold_int = 0 (NN == 1)
                const float v = (float)((double)__fmul_rn(prio, post) / (1.0 + (double)prio));
                const float rt = __fsqrt_rn((float)(0.7212 * (double)v + (double)__fmul_rn(v, v)));
                const float hk = (float)fmax(1.0 / (double)post * (double)rt, 0.001);
                nr.Hk[b] = hk;
                nr.Hk_old[b] = __fmul_rn(__fmul_rn(post, hk), hk);
            }
        }
    }
    __syncwarp();
    if (run2 && lane == 0) {
        nr.init_counter++;
        if (nr.init_counter > 19) { nr.init_counter = 0; nr.first_time = 3; }
    }
    // spectral weighting of the pass-band bins and their mirrors (:2146-2156)
    for (int b = vad_low + lane; b < vad_high; b += 32) {
        const float hk = nr.Hk[b];
        fft[2 * b] = __fmul_rn(fft[2 * b], hk); fft[2 * b + 1] = __fmul_rn(fft[2 * b + 1], hk);
        fft[512 - 2 * b - 2] = __fmul_rn(fft[512 - 2 * b - 2], hk); fft[512 - 2 * b - 1] = __fmul_rn(fft[512 - 2 * b - 1], hk);
    }
    __syncwarp();
    fft_warp_smem<256, 8>(fft, pool + p.tw256_off, true, lane);
    // window on exit + overlap-add (:2165-2189)
    for (int i = lane; i < 128; i += 32) {
        const float a = __fmul_rn(fft[2 * i], __ldg(win + i));
        const float bnext = __fmul_rn(fft[256 + 2 * i], __ldg(win + 128 + i));
        frame[i] = __fadd_rn(a, nr.last_ifft[i]);
        nr.last_ifft[i] = bnext;
    }
    __syncwarp();
}

// Per block: ISR side on lane 0, then the deferred task (warp-wide).  buf: the block's decimated
// audio (n = 8 samples at 12 ksps) in shared memory, processed in place.
__device__ inline void nr_block(const ChanParams &p, NrState &nr, const float *__restrict__ pool, float *buf, int n, float *fft, int lane)
{
    if (lane == 0) {
        int no_dec = n;
        float x[BLK];
        for (int i = 0; i < n; i++) x[i] = buf[i];
        if (p.nr_decim) {
            // DECIMATE_NR: 4 taps, M = 2 (audio_driver.c:195, :649): y[m] = sum_k c[k] s[2m - 3 + k]
            no_dec = n / 2;
            const float *c = pool + p.nr_dec_c;
            float s[3 + BLK];
            for (int i = 0; i < 3; i++) s[i] = nr.dec_hist[i];
            for (int i = 0; i < n; i++) s[3 + i] = x[i];
            for (int m = 0; m < no_dec; m++) {
                float acc = 0.0f;
                for (int k = 0; k < 4; k++) acc = __fadd_rn(acc, __fmul_rn(s[2 * m + k], __ldg(c + k)));
                x[m] = acc;
            }
            for (int i = 0; i < 3; i++) nr.dec_hist[i] = s[n + i];
        }
        float *inb = nr.bufs[nr.fill_in_pt];
        for (int k = 0; k < no_dec; k += 2) {
            inb[2 * nr.trans_count_in] = x[k];
            inb[2 * nr.trans_count_in + 1] = x[k + 1];
            nr.trans_count_in++;
        }
        if (nr.trans_count_in >= 64) {
            const int next = (nr.in_head + 1) % NR_FIFO;
            if (next != nr.in_tail) { nr.in_fifo[nr.in_head] = nr.fill_in_pt; nr.in_head = next; }
            nr.trans_count_in = 0;
            nr.fill_in_pt = (nr.fill_in_pt + 1) % 4;
        }
        if (nr.out_buffer < 0 && nr_fifo_count(nr.out_head, nr.out_tail) > 1) nr.out_buffer = nr.out_fifo[nr.out_tail];
        float dec[BLK];
        if (nr.out_buffer >= 0) {
            const float *ob = nr.bufs[nr.out_buffer] + 128;
            for (int j = 0; j < no_dec; j += 2) {
                dec[j] = ob[2 * nr.outbuff_count];
                dec[j + 1] = ob[2 * nr.outbuff_count + 1];
                nr.outbuff_count++;
            }
            if (nr.outbuff_count >= 64) {
                nr.outbuff_count = 0;
                if (nr.out_head != nr.out_tail) nr.out_tail = (nr.out_tail + 1) % NR_FIFO;
                nr.out_buffer = (nr.out_head != nr.out_tail) ? nr.out_fifo[nr.out_tail] : -1;
            }
        } else {
            for (int j = 0; j < no_dec; j++) dec[j] = 0.0f;
        }
        if (p.nr_decim) {
            // INTERPOLATE_NR: L = 2, 40 taps -> phase length 20 (audio_driver.c:198, :653), then x2.0
            const float *c = pool + p.nr_int_c;
            float s[19 + BLK];
            for (int i = 0; i < 19; i++) s[i] = nr.int_hist[i];
            for (int i = 0; i < no_dec; i++) s[19 + i] = dec[i];
            for (int i = 0; i < no_dec; i++) {
                for (int j = 0; j < 2; j++) {
                    float sum = 0.0f;
                    for (int k = 0; k < 20; k++) sum = __fadd_rn(sum, __fmul_rn(s[i + k], __ldg(c + (1 - j) + 2 * k)));
                    buf[2 * i + j] = __fmul_rn(sum, 2.0f);
                }
            }
            for (int i = 0; i < 19; i++) nr.int_hist[i] = s[no_dec + i];
        } else {
            for (int i = 0; i < n; i++) buf[i] = dec[i];
        }
        // deferred task bookkeeping (AudioNr_HandleNoiseReduction)
        if (!nr.was_here) { nr.was_here = 1; nr.current_buffer_idx = 0; nr.in_tail = nr.in_head; nr.out_tail = nr.out_head; }
    }
    __syncwarp();
    __threadfence_block();
    const int pending = nr_fifo_count(nr.in_head, nr.in_tail) && (NR_FIFO - 1 - nr_fifo_count(nr.out_head, nr.out_tail));
    if (pending) {
        const int cur = nr.current_buffer_idx % 4;
        const int k = nr.in_fifo[nr.in_tail];
        __syncwarp();
        if (p.nb_enable) { nb_frame(p, nr, nr.bufs[k], fft, lane); __threadfence_block(); __syncwarp(); }   // AudioNr_RunNoiseReduction, audio_nr.c:362-365
        if (p.nr_spectral) nr_spectral(p, nr, pool, nr.bufs[k], fft, lane);
        __threadfence_block();
        __syncwarp();
        for (int i = lane; i < 128; i += 32) nr.bufs[cur][128 + i] = nr.bufs[k][i];
        __syncwarp();
        if (lane == 0) {
            nr.in_tail = (nr.in_tail + 1) % NR_FIFO;
            const int next = (nr.out_head + 1) % NR_FIFO;
            if (next != nr.out_tail) { nr.out_fifo[nr.out_head] = cur; nr.out_head = next; }
            nr.current_buffer_idx = cur + 1;
        }
        __threadfence_block();
        __syncwarp();
    }
}

// The same walk for a whole time slice with the warp's lanes over samples (rx_nr_kernel): nr_block above does the interface work --
// 4-tap decimation by 2, packing into the frame buffers, FIFO bookkeeping, unpacking, 40-tap interpolation by 2 -- on lane 0 with the
// bookkeeping in global memory, which costs more than the FFT frames themselves.  Here the bookkeeping scalars live in registers
// (uniform across the warp) for the launch, the two small FIRs run with one output per lane (taps in registers, every sum in the
// reference's order: same bits as nr_block), and the copies are coalesced.  `hs`: 64 floats of shared memory of this warp.
struct NrRun {
    int trans_count_in, outbuff_count, fill_in_pt, out_buffer, in_head, in_tail, out_head, out_tail, current_buffer_idx, was_here;
    int in_fifo[NR_FIFO], out_fifo[NR_FIFO];
};
__device__ __forceinline__ int fifo_get(const int (&f)[NR_FIFO], int i)
{
    int v = f[0];
#pragma unroll
    for (int q = 1; q < NR_FIFO; q++) v = (i == q) ? f[q] : v;
    return v;
}
__device__ __forceinline__ void fifo_set(int (&f)[NR_FIFO], int i, int v)
{
#pragma unroll
    for (int q = 0; q < NR_FIFO; q++) f[q] = (i == q) ? v : f[q];
}

__device__ inline void nr_slice(const ChanParams &p, NrState &nr, const float *__restrict__ pool, float *sc, int nblocks, int n, float *fft,
                                float *hs, int lane)
{
    NrRun r;
    r.trans_count_in = nr.trans_count_in; r.outbuff_count = nr.outbuff_count; r.fill_in_pt = nr.fill_in_pt; r.out_buffer = nr.out_buffer;
    r.in_head = nr.in_head; r.in_tail = nr.in_tail; r.out_head = nr.out_head; r.out_tail = nr.out_tail;
    r.current_buffer_idx = nr.current_buffer_idx; r.was_here = nr.was_here;
#pragma unroll
    for (int q = 0; q < NR_FIFO; q++) { r.in_fifo[q] = nr.in_fifo[q]; r.out_fifo[q] = nr.out_fifo[q]; }
    const bool decim = p.nr_decim != 0;
    const int no_dec = decim ? n / 2 : n;
    const int per = no_dec / 2;                      // sample pairs a block moves in and out of the frame buffers
    // Between two FIFO events (input frame full, output frame used up, a frame processed) the interface is plain streaming: RUNS of
    // up to 32 blocks are decimated, packed, unpacked and interpolated together, lanes over the samples of the run; the events are
    // handled at the last block of a run exactly as the block-by-block walk of nr_block does.
    // staging in the FFT buffer (free between frames): sd = [3 history | run input], si = [19 history | run output frames];
    // the histories themselves live in hs across frames.
    float *sd = fft, *si = fft + 272;
    float *hd = hs, *hi = hs + 4;
    float cd[4], ci[20];
#pragma unroll
    for (int k = 0; k < 4; k++) cd[k] = decim ? __ldg(pool + p.nr_dec_c + k) : 0.0f;
#pragma unroll
    for (int k = 0; k < 20; k++) ci[k] = decim ? __ldg(pool + p.nr_int_c + (1 - (lane & 1)) + 2 * k) : 0.0f;
    if (decim) {
        if (lane < 3) hd[lane] = nr.dec_hist[lane];
        if (lane < 19) hi[lane] = nr.int_hist[lane];
    }
    __syncwarp();
    int blk = 0;
    while (blk < nblocks) {
        // ---- length of the run: up to and including the block of the next event ----
        if (r.out_buffer < 0 && nr_fifo_count(r.out_head, r.out_tail) > 1) r.out_buffer = fifo_get(r.out_fifo, r.out_tail);
        int R = min(nblocks - blk, 32);
        R = min(R, (64 - r.trans_count_in + per - 1) / per);
        if (r.out_buffer >= 0) R = min(R, (64 - r.outbuff_count + per - 1) / per);
        const bool pend_now = nr_fifo_count(r.in_head, r.in_tail) && (NR_FIFO - 1 - nr_fifo_count(r.out_head, r.out_tail));
        if (pend_now || !r.was_here) R = 1;          // a waiting frame is processed after every block; the first block ever resets the FIFOs
        const int n_in = R * n, n_dec = R * no_dec;
        float *buf = sc + (size_t)blk * n;
        // ---- input side: (decimate,) pack ----
        float *inb = nr.bufs[r.fill_in_pt] + 2 * r.trans_count_in;
        if (decim) {
            // DECIMATE_NR: 4 taps, M = 2 (audio_driver.c:195, :649): y[m] = sum_k c[k] s[2m - 3 + k]
            if (lane < 3) sd[lane] = hd[lane];
            for (int i = lane; i < n_in; i += 32) sd[3 + i] = buf[i];
            __syncwarp();
            for (int m = lane; m < n_dec; m += 32) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; k++) acc = __fadd_rn(acc, __fmul_rn(sd[2 * m + k], cd[k]));
                inb[m] = acc;
            }
            if (lane < 3) hd[lane] = sd[n_in + lane];
        } else {
            for (int i = lane; i < n_in; i += 32) inb[i] = buf[i];
        }
        r.trans_count_in += R * per;
        if (r.trans_count_in >= 64) {
            const int next = (r.in_head + 1) % NR_FIFO;
            if (next != r.in_tail) { fifo_set(r.in_fifo, r.in_head, r.fill_in_pt); r.in_head = next; }
            r.trans_count_in = 0;
            r.fill_in_pt = (r.fill_in_pt + 1) % 4;
        }
        // ---- output side: unpack, (interpolate) ----
        const float *ob = (r.out_buffer >= 0) ? nr.bufs[r.out_buffer] + 128 + 2 * r.outbuff_count : nullptr;
        if (decim) {
            // INTERPOLATE_NR: L = 2, 40 taps -> phase length 20 (audio_driver.c:198, :653), then x2.0; output o = 2 i + j
            if (lane < 19) si[lane] = hi[lane];
            for (int j = lane; j < n_dec; j += 32) si[19 + j] = ob ? ob[j] : 0.0f;
            __syncwarp();
            for (int o = lane; o < n_in; o += 32) {
                const int i = o >> 1;
                float sum = 0.0f;
#pragma unroll
                for (int k = 0; k < 20; k++) sum = __fadd_rn(sum, __fmul_rn(si[i + k], ci[k]));
                buf[o] = __fmul_rn(sum, 2.0f);
            }
            if (lane < 19) hi[lane] = si[n_dec + lane];
        } else {
            for (int i = lane; i < n_in; i += 32) buf[i] = ob ? ob[i] : 0.0f;
        }
        if (r.out_buffer >= 0) {
            r.outbuff_count += R * per;
            if (r.outbuff_count >= 64) {
                r.outbuff_count = 0;
                if (r.out_head != r.out_tail) r.out_tail = (r.out_tail + 1) % NR_FIFO;
                r.out_buffer = (r.out_head != r.out_tail) ? fifo_get(r.out_fifo, r.out_tail) : -1;
            }
        }
        __syncwarp();
        // ---- deferred task (AudioNr_HandleNoiseReduction) after the last block of the run ----
        if (!r.was_here) { r.was_here = 1; r.current_buffer_idx = 0; r.in_tail = r.in_head; r.out_tail = r.out_head; }
        const int pending = nr_fifo_count(r.in_head, r.in_tail) && (NR_FIFO - 1 - nr_fifo_count(r.out_head, r.out_tail));
        if (pending) {
            const int cur = r.current_buffer_idx % 4;
            const int k = fifo_get(r.in_fifo, r.in_tail);
            __threadfence_block();
            __syncwarp();
            if (p.nb_enable) { nb_frame(p, nr, nr.bufs[k], fft, lane); __threadfence_block(); __syncwarp(); }   // AudioNr_RunNoiseReduction, audio_nr.c:362-365
            if (p.nr_spectral) nr_spectral(p, nr, pool, nr.bufs[k], fft, lane);
            __threadfence_block();
            __syncwarp();
            for (int i = lane; i < 128; i += 32) nr.bufs[cur][128 + i] = nr.bufs[k][i];
            r.in_tail = (r.in_tail + 1) % NR_FIFO;
            const int next = (r.out_head + 1) % NR_FIFO;
            if (next != r.out_tail) { fifo_set(r.out_fifo, r.out_head, cur); r.out_head = next; }
            r.current_buffer_idx = cur + 1;
            __threadfence_block();
            __syncwarp();
        }
        blk += R;
    }
    __syncwarp();
    if (decim) {
        if (lane < 3) nr.dec_hist[lane] = hd[lane];
        if (lane < 19) nr.int_hist[lane] = hi[lane];
    }
    if (lane == 0) {
        nr.trans_count_in = r.trans_count_in; nr.outbuff_count = r.outbuff_count; nr.fill_in_pt = r.fill_in_pt; nr.out_buffer = r.out_buffer;
        nr.in_head = r.in_head; nr.in_tail = r.in_tail; nr.out_head = r.out_head; nr.out_tail = r.out_tail;
        nr.current_buffer_idx = r.current_buffer_idx; nr.was_here = r.was_here;
#pragma unroll
        for (int q = 0; q < NR_FIFO; q++) { nr.in_fifo[q] = r.in_fifo[q]; nr.out_fifo[q] = r.out_fifo[q]; }
    }
}

}  // namespace uhsdr
