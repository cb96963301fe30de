// rx_front2.cu -- time-parallel half of the split general receiver path, second generation: one warp per channel,
// 512-sample chunks, every FIR of the chain register-blocked (fir_device.cuh: taps staged in shared memory, float4 window
// loads, 4 outputs per lane) instead of one shared-memory load and one global tap load per multiply-add.
//
// Same contract as rx_generic_kernel<true> (rx_generic.cu): sample formatting, clip detection, IQ correction, spectrum /
// zoom-FFT taps, frequency translation (AudioDriver_RxProcessor, mchf-eclipse/drivers/audio/audio_driver.c:2660-2716) and the
// FIR stages of the four topologies (:2718-2829); the result goes to a.scratch for the sample-serial kernels:
//   SSB, decimate first   83-tap /4 decimator on I, Q -> 199-tap Hilbert pair @12 ksps -> I +- Q        [nblocks * 8]
//   SSB, Hilbert first    89-tap Hilbert pair @48 ksps -> I +- Q -> audio decimator /4 or /2            [nblocks * 32 / M]
//   AM / SAM              low-pass decimator /4 or /2 on I, Q (reference summation order in BOTH builds) 2 x [nblocks * 32 / M]
//   FM                    89-tap low-pass pair @48 ksps                                                 2 x [nblocks * 32]
// All sums visit the taps in the reference's order, so the exact build stays bit-identical.
#include "fir_device.cuh"
#include "kernels.h"

namespace uhsdr {

namespace {

constexpr int F2_WARPS = 4;
constexpr int CB = 16;                 // blocks per chunk
constexpr int CS = CB * BLK;           // 512 samples
constexpr int T1MAX = 96, T2MAX = 208; // padded tap counts (stage 1 <= 89 + pad, stage 2 <= 199 + pad)

struct Front2Work {
    alignas(16) float xi[H1 + CS + 8], xq[H1 + CS + 8];   // stage-1 input @48k: [history | new | finite slack for the padded tap tail]
    alignas(16) float bi[H2 + CS + 16], bq[H2 + CS + 16]; // stage-2 input:      [history | new | slack]; the "new" part doubles as the
                                                          // front end's [block][33] staging (16 x 33 = 528 floats)
    alignas(16) float t1i[T1MAX], t1q[T1MAX], t2i[T2MAX], t2q[T2MAX];
    float scr[2 * BLK];
};

__device__ __forceinline__ void shift_hist(float *buf, int H, int nnew, int lane)
{
    // keep the newest H samples: buf[0..H) = buf[nnew..nnew+H).  nnew >= H moves disjoint ranges; otherwise go through registers.
    float tmp[(H2 + 31) / 32];
    int cnt = 0;
    for (int i = lane; i < H; i += 32) tmp[cnt++] = buf[nnew + i];
    __syncwarp();
    cnt = 0;
    for (int i = lane; i < H; i += 32) buf[i] = tmp[cnt++];
    __syncwarp();
}

}  // namespace

__global__ void __launch_bounds__(32 * F2_WARPS, 4)      // 4 CTAs per SM is what the staging buffers allow: keep the registers there too
rx_front2_kernel(RxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * F2_WARPS + warp;
    if (slot >= a.num_items) return;
    const int ch = a.chan_list ? a.chan_list[slot] : slot;
    Front2Work &w = reinterpret_cast<Front2Work *>(smem_raw)[warp];
    const ChanParams &p = a.params[ch];
    ChanState *gst = a.state + ch;
    const float *__restrict__ pool = a.pool;

    // ---- state and taps -> shared memory -------------------------------------------------------
    // (all loads of a group first, then the stores: a launch covers only a few chunks, so the prologue's load latencies count)
    {
        float h1i[H1 / 32], h1q[H1 / 32], h2i[(H2 + 31) / 32], h2q[(H2 + 31) / 32];
#pragma unroll
        for (int k = 0; k < H1 / 32; k++) { h1i[k] = gst->s1_hist_i[lane + 32 * k]; h1q[k] = gst->s1_hist_q[lane + 32 * k]; }
#pragma unroll
        for (int k = 0; k < (H2 + 31) / 32; k++) {
            const int i = lane + 32 * k;
            h2i[k] = (i < H2) ? gst->s2_hist_i[i] : 0.0f; h2q[k] = (i < H2) ? gst->s2_hist_q[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < H1 / 32; k++) { w.xi[lane + 32 * k] = h1i[k]; w.xq[lane + 32 * k] = h1q[k]; }
#pragma unroll
        for (int k = 0; k < (H2 + 31) / 32; k++) {
            const int i = lane + 32 * k;
            if (i < H2) { w.bi[i] = h2i[k]; w.bq[i] = h2q[k]; }
        }
    }
    for (int i = lane; i < CS + 8; i += 32) { w.xi[H1 + i] = 0.0f; w.xq[H1 + i] = 0.0f; }
    for (int i = lane; i < CS + 16; i += 32) { w.bi[H2 + i] = 0.0f; w.bq[H2 + i] = 0.0f; }
    const int N1 = p.s1_ntaps, M1 = p.s1_M, N2 = p.s2_ntaps, M2 = p.s2_M, topo = p.topo;
    fir_stage_taps(w.t1i, pool + p.s1_ci, N1, lane);
    fir_stage_taps(w.t1q, pool + p.s1_cq, N1, lane);
    if (topo == TOPO_SSB_DEC_FIRST) { fir_stage_taps(w.t2i, pool + p.s2_ci, N2, lane); fir_stage_taps(w.t2q, pool + p.s2_cq, N2, lane); }
    else if (topo == TOPO_SSB_HIL_FIRST) fir_stage_taps(w.t2i, pool + p.s2_ci, N2, lane);
    const int g1 = fir_padded_len(N1) / 4, off1 = N1 - 1 + fir_pad_front(N1);     // groups of 4 taps, window reach back from the output's newest sample
    const int g2 = fir_padded_len(N2) / 4, off2 = N2 - 1 + fir_pad_front(N2);
    // front-end state (lane-uniform copies)
    float teta1 = gst->teta1_old, teta2 = gst->teta2_old, teta3 = gst->teta3_old, M_c1 = gst->M_c1, M_c2 = gst->M_c2;
    float osc_q = gst->osc_vect_q, osc_i = gst->osc_vect_i;
    int conv = gst->conversion_freq;
    uint32_t samp_ptr = gst->samp_ptr;
    if (p.shift_kind != 0 && conv != p.shift_freq) { conv = p.shift_freq; osc_i = 0.0f; osc_q = 1.0f; }   // FreqShift re-prepares the NCO, freq_shift.c:289-305
    __syncwarp();

    // per-channel constants the block loops use, in registers (the loops store to global memory, so the compiler would re-read
    // them from the parameter block every iteration: a global-load latency per block)
    const int k_iq_auto = p.iq_auto, k_zoom_m = p.zoom_m, k_shift_kind = p.shift_kind, k_shift_down = p.shift_down, k_lsb = p.lsb;
    const float k_adj_i = p.adj_i, k_adj_q = p.adj_q, k_phase_bal = p.phase_bal, k_osc_cos = p.osc_cos, k_osc_sin = p.osc_sin;
    const int M = p.M;
    const int ndec_blk = BLK / M;
    float *spec_ring = (p.spectrum_enable && a.spec_ring) ? (a.spec_ring + (size_t)ch * 1024) : nullptr;
    const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
    const int2 *__restrict__ iq = reinterpret_cast<const int2 *>(a.iq) + chan_base;
    float *sc = a.scratch + (size_t)slot * (size_t)a.scratch_stride;
    int clip_q = 0, clip_h = 0, clip_f = 0;

    for (int blk0 = 0; blk0 < a.nblocks; blk0 += CB) {
        const int nb = min(CB, a.nblocks - blk0);
        const int ns = nb * BLK;
        const int ndec = nb * ndec_blk;

        // ---- front end: format, IQ correction, spectrum tap, frequency translation ----
        // The reference walks block by block (statistics of the block -> low-pass -> M_c1 / M_c2 -> correction of the same block);
        // done that way on the GPU every block pays a global-load latency and a chain of three IEEE divisions and a square root.
        // Here the chunk's blocks are (1) loaded and converted together, (2) summed with lane b on block b in sample order (the
        // reference's order, so both builds keep its sums), (3) low-passed block after block (the only sequential part: three
        // double multiply-adds), (4) turned into M_c1 / M_c2 with lane b on block b, (5) corrected block by block.
        float *ti = w.bi + H2, *tq = w.bq + H2;          // staging [block][33]: free until the FIR stages write their outputs there
        {
            int2 raw[CB];
#pragma unroll
            for (int b = 0; b < CB; b++) raw[b] = (b < nb) ? iq[(size_t)(blk0 + b) * BLK + lane] : make_int2(0, 0);
#pragma unroll
            for (int b = 0; b < CB; b++) {
                const int level = abs(raw[b].x) >> 16;                            // audio_driver.c:2660-2685
                clip_q |= (level > 4096 / 4); clip_h |= (level > 4096 / 2); clip_f |= (level > 4096);
                ti[b * 33 + lane] = __fmul_rn((float)raw[b].x, 0.0000152587890625f);
                tq[b * 33 + lane] = __fmul_rn((float)raw[b].y, 0.0000152587890625f);
            }
        }
        __syncwarp();
        float mc1 = M_c1, mc2 = M_c2;                    // lane b: the factors of block b
        if (k_iq_auto) {
            // audio_driver.c:2274-2313 (Moseley & Slump): block statistics, EMA in double
            float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            if (lane < nb) {
                const float *ri = ti + lane * 33, *rq = tq + lane * 33;
#pragma unroll 8
                for (int j = 0; j < BLK; j++) {
                    const float vi = ri[j], vq = rq[j];
                    s1 = __fadd_rn(s1, __fmul_rn(sign_new(vi), vq));
                    s2 = __fadd_rn(s2, __fmul_rn(sign_new(vi), vi));
                    s3 = __fadd_rn(s3, __fmul_rn(sign_new(vq), vq));
                }
            }
            float m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
            for (int b = 0; b < nb; b++) {
                const float b1 = __shfl_sync(0xffffffffu, s1, b), b2 = __shfl_sync(0xffffffffu, s2, b), b3 = __shfl_sync(0xffffffffu, s3, b);
                teta1 = (float)(-0.003 * (double)__fdiv_rn(b1, 32.0f) + 0.997 * (double)teta1);
                teta2 = (float)(0.003 * (double)__fdiv_rn(b2, 32.0f) + 0.997 * (double)teta2);
                teta3 = (float)(0.003 * (double)__fdiv_rn(b3, 32.0f) + 0.997 * (double)teta3);
                if (lane == b) { m1 = teta1; m2 = teta2; m3 = teta3; }
            }
            mc1 = (m2 != 0.0f) ? __fdiv_rn(m1, m2) : 0.0f;
            float help = __fmul_rn(m2, m2);
            if (help > 0.0f) help = __fdiv_rn(__fsub_rn(__fmul_rn(m3, m3), __fmul_rn(m1, m1)), help);
            mc2 = (help > 0.0f) ? __fsqrt_rn(help) : 1.0f;
            M_c1 = __shfl_sync(0xffffffffu, mc1, nb - 1); M_c2 = __shfl_sync(0xffffffffu, mc2, nb - 1);
        }
        for (int b = 0; b < nb; b++) {
            float fi = ti[b * 33 + lane], fq = tq[b * 33 + lane];
            if (k_iq_auto) {
                const float c1 = __shfl_sync(0xffffffffu, mc1, b), c2 = __shfl_sync(0xffffffffu, mc2, b);
                fq = __fadd_rn(fq, __fmul_rn(c1, fi));
                fi = __fmul_rn(fi, c2);
            } else {
                fi = __fmul_rn(fi, k_adj_i);                                     // manual gain + phase, audio_driver.c:2259-2267
                fq = __fmul_rn(fq, k_adj_q);
                if (k_phase_bal < 0.0f) fq = __fadd_rn(fq, __fmul_rn(fi, k_phase_bal));
                else if (k_phase_bal > 0.0f) fi = __fadd_rn(fi, __fmul_rn(fq, k_phase_bal));
            }
            if (spec_ring && k_zoom_m == 0) {                                    // AudioDriver_SpectrumNoZoomProcessSamples, :1811-1849
                uint32_t ptr = samp_ptr + 2u * (uint32_t)lane;
                if (ptr >= 1024u) ptr -= 1024u;
                spec_ring[ptr] = fq; spec_ring[ptr + 1] = fi;
                samp_ptr += 64u; if (samp_ptr >= 1024u) samp_ptr -= 1024u;
            }
            if (k_shift_kind == 1) {                                             // FreqShift_QuarterFs, freq_shift.c:219-262
                float ib = k_shift_down ? fq : fi, qb = k_shift_down ? fi : fq;
                const int ph = lane & 3;
                float ni = ib, nq = qb;
                if (ph == 1) { ni = qb; nq = -ib; }
                else if (ph == 2) { ni = -ib; nq = -qb; }
                else if (ph == 3) { ni = -qb; nq = ib; }
                if (k_shift_down) { fq = ni; fi = nq; } else { fi = ni; fq = nq; }
            } else if (k_shift_kind == 2) {                                      // FreqShift_Approx, freq_shift.c:57-108
                if (lane == 0) {
                    float vq = osc_q, vi = osc_i;
                    for (int n = 0; n < BLK; n++) {
                        const float oq = __fsub_rn(__fmul_rn(vq, k_osc_cos), __fmul_rn(vi, k_osc_sin));
                        const float oi = __fadd_rn(__fmul_rn(vi, k_osc_cos), __fmul_rn(vq, k_osc_sin));
                        w.scr[n] = oq; w.scr[BLK + n] = oi;
                        vq = oq; vi = oi;
                    }
                    const float g = __fdiv_rn(__fsub_rn(3.0f, __fadd_rn(__fmul_rn(vq, vq), __fmul_rn(vi, vi))), 2.0f);
                    osc_q = __fmul_rn(g, vq); osc_i = __fmul_rn(g, vi);
                }
                osc_q = __shfl_sync(0xffffffffu, osc_q, 0); osc_i = __shfl_sync(0xffffffffu, osc_i, 0);
                __syncwarp();
                const float oq = w.scr[lane], oi = w.scr[BLK + lane];
                float ib = k_shift_down ? fq : fi, qb = k_shift_down ? fi : fq;
                const float nq = __fsub_rn(__fmul_rn(qb, oq), __fmul_rn(ib, oi));
                const float ni = __fadd_rn(__fmul_rn(ib, oq), __fmul_rn(qb, oi));
                if (k_shift_down) { fq = ni; fi = nq; } else { fi = ni; fq = nq; }
                __syncwarp();
            }
            if (spec_ring && k_zoom_m != 0) {
                // AudioDriver_SpectrumZoomProcessSamples, :1860-1909: 4-stage DF1 low-pass on I and on Q (lane 0 / lane 1), 4-tap
                // decimation by 2^magnify, 32 >> magnify (Q, I) pairs into the ring; reference operation order
                __syncwarp();
                w.scr[lane] = fi; w.scr[BLK + lane] = fq;
                __syncwarp();
                const int nout = BLK >> k_zoom_m, MZ = 1 << k_zoom_m;
                if (lane < 2) {
                    float *buf = w.scr + lane * BLK;
                    BiquadS *bs = lane ? gst->zoom_bq_q : gst->zoom_bq_i;
                    const float *zc = pool + p.zoom_bq_off;
                    for (int sg = 0; sg < 4; sg++) {
                        const float b0 = __ldg(zc + 5 * sg), b1 = __ldg(zc + 5 * sg + 1), b2 = __ldg(zc + 5 * sg + 2), a1 = __ldg(zc + 5 * sg + 3), a2 = __ldg(zc + 5 * sg + 4);
                        BiquadS s_ = bs[sg];
                        for (int i = 0; i < BLK; i++) {
                            const float x = buf[i];
                            const float acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(b0, x), __fmul_rn(b1, s_.x1)), __fmul_rn(b2, s_.x2)), __fmul_rn(a1, s_.y1)), __fmul_rn(a2, s_.y2));
                            s_.x2 = s_.x1; s_.x1 = x; s_.y2 = s_.y1; s_.y1 = acc;
                            buf[i] = acc;
                        }
                        bs[sg] = s_;
                    }
                    float *hist = lane ? gst->zoom_hist_q : gst->zoom_hist_i;
                    const float *dc = pool + p.zoom_dec_off;
                    const float c0 = __ldg(dc), c1 = __ldg(dc + 1), c2 = __ldg(dc + 2), c3 = __ldg(dc + 3);
                    const float h0 = hist[0], h1 = hist[1], h2 = hist[2];
                    hist[0] = buf[BLK - 3]; hist[1] = buf[BLK - 2]; hist[2] = buf[BLK - 1];
                    float outv[BLK / 2];
                    for (int j = 0; j < nout; j++) {
                        const int o = j * MZ;
                        const float s0 = o >= 3 ? buf[o - 3] : (o == 0 ? h0 : (o == 1 ? h1 : h2));
                        const float s1_ = o >= 2 ? buf[o - 2] : (o == 0 ? h1 : h2);
                        const float s2_ = o >= 1 ? buf[o - 1] : h2;
                        const float s3_ = buf[o];
                        outv[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s0, c0), __fmul_rn(s1_, c1)), __fmul_rn(s2_, c2)), __fmul_rn(s3_, c3));
                    }
                    for (int j = 0; j < nout; j++) buf[j] = outv[j];
                }
                __syncwarp();
                if (lane < nout) {
                    uint32_t ptr = samp_ptr + 2u * (uint32_t)lane;
                    if (ptr >= 1024u) ptr -= 1024u;
                    spec_ring[ptr] = w.scr[BLK + lane]; spec_ring[ptr + 1] = w.scr[lane];
                }
                samp_ptr += 2u * (uint32_t)nout; if (samp_ptr >= 1024u) samp_ptr -= 1024u;
                __syncwarp();
            }
            w.xi[H1 + b * BLK + lane] = fi;
            w.xq[H1 + b * BLK + lane] = fq;
        }
        __syncwarp();

        // ---- FIR stages ------------------------------------------------------------------------
        if (topo == TOPO_SSB_DEC_FIRST || topo == TOPO_AM_SAM) {
            // stage 1: decimating FIR pair on I and Q
            if (M1 == 4) {
                float yi[4], yq[4];
                const float *xi0 = w.xi + H1 + 4 * lane - off1, *xq0 = w.xq + H1 + 4 * lane - off1;
                if (topo == TOPO_AM_SAM) { fir_dec4<true, 4>(xi0, w.t1i, g1, yi); fir_dec4<true, 4>(xq0, w.t1q, g1, yq); }
                else { fir_dec4<false, 4>(xi0, w.t1i, g1, yi); fir_dec4<false, 4>(xq0, w.t1q, g1, yq); }
#pragma unroll
                for (int r = 0; r < 4; r++) { w.bi[H2 + lane + 32 * r] = yi[r]; w.bq[H2 + lane + 32 * r] = yq[r]; }
            } else {
                const int np = 4 * g1;
                for (int m = lane; m < ns / M1; m += 32) {
                    const float *xi0 = w.xi + H1 + m * M1 - off1, *xq0 = w.xq + H1 + m * M1 - off1;
                    if (topo == TOPO_AM_SAM) { w.bi[H2 + m] = fir_scalar<true>(xi0, w.t1i, np); w.bq[H2 + m] = fir_scalar<true>(xq0, w.t1q, np); }
                    else { w.bi[H2 + m] = fir_scalar<false>(xi0, w.t1i, np); w.bq[H2 + m] = fir_scalar<false>(xq0, w.t1q, np); }
                }
            }
            __syncwarp();
            if (topo == TOPO_AM_SAM) {
                const size_t half = (size_t)a.nblocks * ndec_blk, o = (size_t)blk0 * ndec_blk;
                for (int m = lane; m < ndec; m += 32) { sc[o + m] = w.bi[H2 + m]; sc[half + o + m] = w.bq[H2 + m]; }
            } else {
                // stage 2: Hilbert pair at the decimated rate, USB = I + Q / LSB = I - Q (:2751-2790); four outputs per lane
                float yi[4], yq[4];
                fir4_m1_pair<false>(w.bi + H2 + 4 * lane - off2, w.t2i, w.bq + H2 + 4 * lane - off2, w.t2q, g2, yi, yq);
                const size_t o = (size_t)blk0 * ndec_blk;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int m = 4 * lane + r;
                    if (m < ndec) sc[o + m] = k_lsb ? __fsub_rn(yi[r], yq[r]) : __fadd_rn(yi[r], yq[r]);
                }
                __syncwarp();
                shift_hist(w.bi, H2, ndec, lane);
                shift_hist(w.bq, H2, ndec, lane);
            }
        } else {
            // stage 1: FIR pair @48k, four consecutive outputs per lane and 128-sample sub-chunk
            const size_t half = (size_t)a.nblocks * BLK, o = (size_t)blk0 * BLK;
            for (int sub = 0; sub < ns; sub += 128) {
                float yi[4], yq[4];
                fir4_m1_pair<false>(w.xi + H1 + sub + 4 * lane - off1, w.t1i, w.xq + H1 + sub + 4 * lane - off1, w.t1q, g1, yi, yq);
                if (topo == TOPO_FM) {
                    if (sub + 4 * lane < ns) {
                        *reinterpret_cast<float4 *>(sc + o + sub + 4 * lane) = make_float4(yi[0], yi[1], yi[2], yi[3]);
                        *reinterpret_cast<float4 *>(sc + half + o + sub + 4 * lane) = make_float4(yq[0], yq[1], yq[2], yq[3]);
                    }
                } else {
                    float au[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) au[r] = k_lsb ? __fsub_rn(yi[r], yq[r]) : __fadd_rn(yi[r], yq[r]);      // combine at 48k (:2781-2803)
                    *reinterpret_cast<float4 *>(w.bi + H2 + sub + 4 * lane) = make_float4(au[0], au[1], au[2], au[3]);
                }
            }
            __syncwarp();
            if (topo == TOPO_SSB_HIL_FIRST) {
                // stage 2: decimate the audio
                const size_t od = (size_t)blk0 * ndec_blk;
                if (M2 == 4) {
                    float y[4];
                    fir_dec4<false, 4>(w.bi + H2 + 4 * lane - off2, w.t2i, g2, y);
#pragma unroll
                    for (int r = 0; r < 4; r++) if (lane + 32 * r < ndec) sc[od + lane + 32 * r] = y[r];
                } else {
                    const int np = 4 * g2;
                    for (int m = lane; m < ndec; m += 32) sc[od + m] = fir_scalar<false>(w.bi + H2 + m * M2 - off2, w.t2i, np);
                }
                __syncwarp();
                shift_hist(w.bi, H2, ns, lane);
            }
        }
        __syncwarp();
        shift_hist(w.xi, H1, ns, lane);
        shift_hist(w.xq, H1, ns, lane);
    }

    // ---- store the state this kernel owns ---------------------------------------------------
    clip_q = __any_sync(0xffffffffu, clip_q); clip_h = __any_sync(0xffffffffu, clip_h); clip_f = __any_sync(0xffffffffu, clip_f);
    for (int i = lane; i < H1; i += 32) { gst->s1_hist_i[i] = w.xi[i]; gst->s1_hist_q[i] = w.xq[i]; }
    for (int i = lane; i < H2; i += 32) { gst->s2_hist_i[i] = w.bi[i]; gst->s2_hist_q[i] = w.bq[i]; }
    if (lane == 0) {
        gst->teta1_old = teta1; gst->teta2_old = teta2; gst->teta3_old = teta3; gst->M_c1 = M_c1; gst->M_c2 = M_c2;
        gst->osc_vect_q = osc_q; gst->osc_vect_i = osc_i; gst->conversion_freq = conv;
        gst->samp_ptr = samp_ptr;
        if (clip_q) gst->adc_quarter_clip = 1;
        if (clip_h) gst->adc_half_clip = 1;
        if (clip_f) gst->adc_clip = 1;
        gst->blocks += a.nblocks;
    }
}

// every shape the kernel's buffers and the 16-byte window alignment can take; anything else stays on rx_generic_kernel<true>
bool rx_front2_eligible(const ChanParams &p)
{
    if (p.topo < TOPO_SSB_DEC_FIRST || p.topo > TOPO_FM) return false;
    if (p.s1_ntaps < 1 || fir_padded_len(p.s1_ntaps) > T1MAX || p.s1_ntaps - 1 + fir_pad_front(p.s1_ntaps) > H1) return false;
    if (p.topo == TOPO_SSB_DEC_FIRST || p.topo == TOPO_SSB_HIL_FIRST) {
        if (p.s2_ntaps < 1 || fir_padded_len(p.s2_ntaps) > T2MAX || p.s2_ntaps - 1 + fir_pad_front(p.s2_ntaps) > H2) return false;
    }
    if (p.topo == TOPO_SSB_DEC_FIRST) return p.s1_M == 4 && p.s2_M == 1 && p.M == 4;
    if (p.topo == TOPO_AM_SAM) return (p.s1_M == 4 || p.s1_M == 2) && p.M == p.s1_M;
    if (p.topo == TOPO_SSB_HIL_FIRST) return p.s1_M == 1 && (p.s2_M == 4 || p.s2_M == 2) && p.M == p.s2_M;
    return p.s1_M == 1;      // FM
}

cudaError_t launch_rx_front2(const RxArgs &a, cudaStream_t stream)
{
    if (a.scratch == nullptr || a.chan_list == nullptr) return cudaErrorInvalidValue;
    const size_t smem = sizeof(Front2Work) * F2_WARPS;
    cudaError_t e = cudaFuncSetAttribute(rx_front2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (a.num_items + F2_WARPS - 1) / F2_WARPS;
    if (grid == 0) return cudaSuccess;
    rx_front2_kernel<<<grid, 32 * F2_WARPS, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace uhsdr
