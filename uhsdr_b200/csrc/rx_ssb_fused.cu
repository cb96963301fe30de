// rx_ssb_fused.cu -- fused narrow-SSB/CW receiver kernel (BASELINE.json configs[1]).
//
// One persistent CTA owns up to 32 channels for the whole launch and every sample crosses HBM
// once: int32 I/Q in, int32 audio out.  The reference chain for these filter paths
// (FilterPathInfo[4..47], mchf-eclipse/drivers/audio/audio_filter.c:147-922; flow in
// audio_driver.c:2603-2942) is
//
//   format + IQ correction + Fs/4 translate -> 83-tap /4 decimator on I and Q
//   -> 199-tap Hilbert pair @12 ksps -> I +/- Q -> 10-stage lattice IIR -> WDSP AGC -> gain
//   -> 4-stage biquad -> x4 polyphase interpolator -> (6-stage anti-alias lattice) -> treble
//   biquad -> x10 -> int32 << 16.
//
// The time-parallel part (front end + both FIR pairs, 282 of the 346 FLOP per sample) runs on
// "FIR warps": 8 lanes per channel, 4 decimated outputs per lane, coefficients as immediate
// constant-bank operands (the taps travel as a __grid_constant__ kernel parameter), samples staged
// in shared memory (polyphase layout for the decimator, so every LDS is a conflict-free 128-bit
// load) and streamed once through 4 FMAs each.  The sample-serial recurrences run one channel per
// lane on four specialised warps (lattice | AGC | EQ + interpolator | anti-alias + treble) that
// form a software pipeline with the FIR warps through double-buffered, channel-minor shared-memory
// queues; one __syncthreads per 128-sample chunk advances the pipeline.  The AGC's sliding
// 49-sample maximum (audio_agc.c:409-429 rescans on demand, divergent) is computed with the
// van Herk / Gil-Werman block decomposition, which yields the same exact maximum with uniform
// control flow.
#include "dsp_device.cuh"
#include "kernels.h"
#include "uhsdr_b200.h"

namespace uhsdr {

namespace {

constexpr int FG = 28;             // channel slots per CTA (7 FIR warps x 4 channels)
constexpr int CH4 = 128;           // input samples per chunk (4 blocks)
constexpr int ND = 32;             // decimated samples per chunk
constexpr int XP = 56;             // per-phase slots: 24 history (21 used) + 32 new
constexpr int XH = 24;             // history slots per phase (multiple of 4: the slide is 6 float4)
constexpr int XCH = 2 * 4 * XP + 4;   // floats per channel (I phases, Q phases) + 4 -> bank stagger
constexpr int DL = 200 + ND;       // Hilbert input: 200 history slots + 32 new
constexpr int DCH = 2 * DL + 4;     // + 4 -> neighbouring channels land on disjoint banks
constexpr int SMS = 33;            // channel-minor stride of the pipeline queues
constexpr int AGC_W = 49;          // attack_buffsize at 12 ksps (audio_agc.c:290)
constexpr int RING = 64;            // AGC delay ring slots (>= 49 + one group of 8)
constexpr int NWARP_FIR = FG / 4;   // one FIR warp = 4 channels x 8 lanes, 4 decimated outputs per lane
constexpr int DEC_PAD = 32;        // zero padding in front of the decimator taps (FusedCoefs::dec)
constexpr int HIL_PAD = 12;        // zero padding in front of the Hilbert taps
constexpr int W_LAT = NWARP_FIR, W_AGC = NWARP_FIR + 1, W_EQ = NWARP_FIR + 2, W_POST = NWARP_FIR + 3;
constexpr int NTHREADS = 32 * (NWARP_FIR + 4);
constexpr int PIPE_DEPTH = 5;      // FIR t | LAT t-1 | AGC t-2 | EQ t-3 | POST t-4 | WRITE t-5

struct Smem {
    alignas(16) float x[FG * XCH];
    alignas(16) float d[FG * DCH];
    float aud[2][ND * SMS];
    float lat[2][ND * SMS];
    float agc[2][ND * SMS];
    float out[3][CH4 * SMS];
    float ring[RING * SMS];
    float smax[2][ND * SMS];            // suffix maxima of |x| over the previous two 32-sample chunks
    alignas(16) int2 raw[FG * CH4];     // staged input chunk: bulk-async copy target, 1 KB per channel
    alignas(8) unsigned long long mbar[NWARP_FIR];   // one transaction barrier per FIR warp
};

// ---- bulk asynchronous copy (TMA, 1-D) + mbarrier, sm_90+/sm_100 PTX --------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int N> struct IC { static constexpr int value = N; };
template <int I, int N, typename F> __device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) { f(IC<I>{}); static_for<I + 1, N>(f); }
}

__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// ---------------------------------------------------------------------------------------------
// FIR warps
// ---------------------------------------------------------------------------------------------
struct FirLaneState {
    float te1, te2, te3;     // teta*_old
    float c1, c2;            // M_c1, M_c2
    int clip;                // bit0 quarter, bit1 half, bit2 full
};

// Decimator: y[m] = sum_k c[k] x[4m - 82 + k] (arm_fir_decimate_f32.c:455-486).  With 96 history
// slots in front (24 per phase), buffer position b = 4m + k + 14; phase = b & 3, idx = b >> 2 =
// m + 3 + ((k + 2) >> 2).  Each lane makes 4 consecutive outputs m0..m0+3 for I and for Q.
// Elements are consumed in ascending b, which is ascending k for every output (the reference's
// summation order).  The loop is rolled to keep the instruction-cache footprint small; taps outside
// [0, 82] hit the zero padding of FusedCoefs::dec and add +-0.
__device__ __forceinline__ void decimate4(const float *xpi, const float *xpq, int m0, const FusedCoefs &fc,
                                          float ai[4], float aq[4])
{
#pragma unroll
    for (int j = 0; j < 4; j++) { ai[j] = 0.0f; aq[j] = 0.0f; }
#pragma unroll 1
    for (int q = 0; q < 7; q++) {
        float4 vi[4], vq[4];
#pragma unroll
        for (int ph = 0; ph < 4; ph++) { vi[ph] = lds128(xpi + ph * XP + m0 + 4 * q); vq[ph] = lds128(xpq + ph * XP + m0 + 4 * q); }
        const float *cq = fc.dec + DEC_PAD + 16 * q - 14;     // tap index K = 16q + 4(E - 3 - J) + PH - 2
#pragma unroll
        for (int e = 0; e < 4; e++) {
#pragma unroll
            for (int ph = 0; ph < 4; ph++) {
                const float xi = (e == 0) ? vi[ph].x : (e == 1) ? vi[ph].y : (e == 2) ? vi[ph].z : vi[ph].w;
                const float xq = (e == 0) ? vq[ph].x : (e == 1) ? vq[ph].y : (e == 2) ? vq[ph].z : vq[ph].w;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float c = cq[4 * (e - j) + ph];
                    ai[j] = mad(c, xi, ai[j]);
                    aq[j] = mad(c, xq, aq[j]);
                }
            }
        }
    }
}

// Hilbert pair at 12 ksps: y[n] = sum_k c[k] d[n - 198 + k] (arm_fir_f32.c:522-529).  With 200
// history slots, position = n + k + 2; 4 outputs per lane, elements streamed in ascending position.
__device__ __forceinline__ void hilbert4(const float *dpi, const float *dpq, int n0, const FusedCoefs &fc,
                                         float hi[4], float hq[4])
{
#pragma unroll
    for (int j = 0; j < 4; j++) { hi[j] = 0.0f; hq[j] = 0.0f; }
#pragma unroll 3
    for (int q = 0; q < 51; q++) {
        const float4 vi = lds128(dpi + n0 + 4 * q), vq = lds128(dpq + n0 + 4 * q);
        const float *ci = fc.hil_i + HIL_PAD + 4 * q - 2;     // tap index K = 4q + E - J - 2
        const float *cq = fc.hil_q + HIL_PAD + 4 * q - 2;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float xi = (e == 0) ? vi.x : (e == 1) ? vi.y : (e == 2) ? vi.z : vi.w;
            const float xq = (e == 0) ? vq.x : (e == 1) ? vq.y : (e == 2) ? vq.z : vq.w;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                hi[j] = mad(ci[e - j], xi, hi[j]);
                hq[j] = mad(cq[e - j], xq, hq[j]);
            }
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(NTHREADS, 1)
rx_ssb_fused_kernel(const __grid_constant__ RxArgs a, const __grid_constant__ FusedCoefs fc, int chans_per_cta)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cta_first = blockIdx.x * chans_per_cta;
    const int n_here = min(chans_per_cta, a.num_items - cta_first);
    const int nchunks = a.nblocks / 4;
    const float *__restrict__ pool = a.pool;

    if (warp < NWARP_FIR) {
        // ======================= FIR warp: 4 channels x 8 lanes ==================================
        const int cl = lane >> 3, r = lane & 7;
        const int g = warp * 4 + cl;                   // channel slot in the CTA
        const bool active = g < n_here;
        const int ch = active ? a.chan_list[cta_first + g] : a.chan_list[cta_first];
        const ChanParams &p = a.params[ch];
        ChanState *st = a.state + ch;
        float *xi = sm.x + g * XCH, *xq = xi + 4 * XP;
        float *di = sm.d + g * DCH, *dq = di + DL;
        const unsigned gmask = 0xffu << (8 * cl);
        const int n_warp_ch = max(0, min(4, n_here - warp * 4));     // active channels of this warp
        unsigned long long *bar = &sm.mbar[warp];
        const int2 *raw = sm.raw + g * CH4;

        // ---- load histories and IQ-correction state ----
        FirLaneState ls;
        ls.te1 = st->teta1_old; ls.te2 = st->teta2_old; ls.te3 = st->teta3_old; ls.c1 = st->M_c1; ls.c2 = st->M_c2; ls.clip = 0;
        // s1_hist[H1=96]: sample s (-96..-1) at [96 + s] = position b -> phase b&3, idx b>>2
        for (int b = r; b < 4 * XH; b += 8) {
            xi[(b & 3) * XP + (b >> 2)] = active ? st->s1_hist_i[b] : 0.0f;
            xq[(b & 3) * XP + (b >> 2)] = active ? st->s1_hist_q[b] : 0.0f;
        }
        for (int i = r; i < 200; i += 8) {
            di[i] = active ? st->s2_hist_i[i] : 0.0f;
            dq[i] = active ? st->s2_hist_q[i] : 0.0f;
        }
        const int iq_auto = p.iq_auto, shift_kind = p.shift_kind, shift_down = p.shift_down, lsb = p.lsb;
        const float adj_i = p.adj_i, adj_q = p.adj_q, phase_bal = p.phase_bal;
        const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
        const int2 *__restrict__ src = reinterpret_cast<const int2 *>(a.iq) + chan_base;
        int4 *__restrict__ dst = reinterpret_cast<int4 *>(reinterpret_cast<int2 *>(a.audio) + chan_base);
        float2 *__restrict__ dst_f = a.audio_f ? reinterpret_cast<float2 *>(a.audio_f + chan_base) : nullptr;
        const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;

        // ---- stage chunk 0: one bulk asynchronous copy of 1 KB per channel, tracked by the warp's mbarrier ----
        if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        if (nchunks > 0 && n_warp_ch > 0) {
            if (lane == 0) mbar_expect_tx(bar, (unsigned)(n_warp_ch * CH4 * sizeof(int2)));
            __syncwarp();
            if (r == 0 && active) bulk_g2s(sm.raw + g * CH4, src, CH4 * sizeof(int2), bar);
        }
        __syncwarp();

        for (int t = 0; t < nchunks + PIPE_DEPTH; t++) {
            if (t < nchunks) {
                if (n_warp_ch > 0) mbar_wait(bar, (unsigned)(t & 1));
#if UHSDR_EXACT
                // ---- front end, one 32-sample block at a time: pairs p = r + 8i (samples 2p, 2p+1) ----
#pragma unroll 1
                for (int b = 0; b < 4; b++) {
                    float fi[4], fq[4];
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const int4 v = *reinterpret_cast<const int4 *>(raw + b * 32 + 2 * (r + 8 * i));
                        const int lv = max(abs(v.x) >> 16, abs(v.z) >> 16);     // audio_driver.c:2662-2675
                        ls.clip |= (lv > 1024 ? 1 : 0) | (lv > 2048 ? 2 : 0) | (lv > 4096 ? 4 : 0);
                        fi[2 * i] = __fmul_rn((float)v.x, 0.0000152587890625f);
                        fq[2 * i] = __fmul_rn((float)v.y, 0.0000152587890625f);
                        fi[2 * i + 1] = __fmul_rn((float)v.z, 0.0000152587890625f);
                        fq[2 * i + 1] = __fmul_rn((float)v.w, 0.0000152587890625f);
                    }
                    if (iq_auto) {
                        // Moseley & Slump statistics of the block (audio_driver.c:2274-2279)
                        float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#if UHSDR_EXACT
                        // reference order: sample n lives in lane (n/2)&7, slot 2*((n/2)>>3) + (n&1)
#pragma unroll
                        for (int n = 0; n < 32; n++) {
                            const int slot = 2 * ((n >> 1) >> 3) + (n & 1);
                            const float vi = __shfl_sync(gmask, fi[slot], (n >> 1) & 7, 8);
                            const float vq = __shfl_sync(gmask, fq[slot], (n >> 1) & 7, 8);
                            s1 = __fadd_rn(s1, __fmul_rn(sign_new(vi), vq));
                            s2 = __fadd_rn(s2, __fmul_rn(sign_new(vi), vi));
                            s3 = __fadd_rn(s3, __fmul_rn(sign_new(vq), vq));
                        }
#else
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            s1 += __fmul_rn(sign_new(fi[k]), fq[k]);
                            s2 += fabsf(fi[k]);
                            s3 += fabsf(fq[k]);
                        }
#pragma unroll
                        for (int dlt = 1; dlt < 8; dlt <<= 1) {
                            s1 += __shfl_xor_sync(gmask, s1, dlt, 8);
                            s2 += __shfl_xor_sync(gmask, s2, dlt, 8);
                            s3 += __shfl_xor_sync(gmask, s3, dlt, 8);
                        }
#endif
                        // teta = -/+0.003*(sum/32) + 0.997*teta_old, evaluated in double (:2281-2283)
                        ls.te1 = (float)(-0.003 * (double)__fdiv_rn(s1, 32.0f) + 0.997 * (double)ls.te1);
                        ls.te2 = (float)(0.003 * (double)__fdiv_rn(s2, 32.0f) + 0.997 * (double)ls.te2);
                        ls.te3 = (float)(0.003 * (double)__fdiv_rn(s3, 32.0f) + 0.997 * (double)ls.te3);
                        ls.c1 = (ls.te2 != 0.0f) ? __fdiv_rn(ls.te1, ls.te2) : 0.0f;
                        float help = __fmul_rn(ls.te2, ls.te2);
                        if (help > 0.0f) help = __fdiv_rn(__fsub_rn(__fmul_rn(ls.te3, ls.te3), __fmul_rn(ls.te1, ls.te1)), help);
                        ls.c2 = (help > 0.0f) ? __fsqrt_rn(help) : 1.0f;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            fq[k] = __fadd_rn(fq[k], __fmul_rn(ls.c1, fi[k]));
                            fi[k] = __fmul_rn(fi[k], ls.c2);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            float vi = __fmul_rn(fi[k], adj_i), vq = __fmul_rn(fq[k], adj_q);
                            if (phase_bal < 0.0f) vq = __fadd_rn(vq, __fmul_rn(vi, phase_bal));
                            else if (phase_bal > 0.0f) vi = __fadd_rn(vi, __fmul_rn(vq, phase_bal));
                            fi[k] = vi; fq[k] = vq;
                        }
                    }
                    // ---- Fs/4 translate (freq_shift.c:219-262) + store in polyphase layout ----
#pragma unroll
                    for (int i = 0; i < 2; i++) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int n = 32 * b + 2 * (r + 8 * i) + e;      // sample index in the chunk
                            float vi = fi[2 * i + e], vq = fq[2 * i + e];
                            if (shift_kind == 1) {
                                float ib = shift_down ? vq : vi, qb = shift_down ? vi : vq;
                                const int ph = n & 3;
                                float ni = ib, nq = qb;
                                if (ph == 1) { ni = qb; nq = -ib; }
                                else if (ph == 2) { ni = -ib; nq = -qb; }
                                else if (ph == 3) { ni = -qb; nq = ib; }
                                if (shift_down) { vq = ni; vi = nq; } else { vi = ni; vq = nq; }
                            }
                            const int pos = n + 4 * XH;
                            xi[(pos & 3) * XP + (pos >> 2)] = vi;
                            xq[(pos & 3) * XP + (pos >> 2)] = vq;
                        }
                    }
                }
#else
                // ---- front end (shipping build): the whole 128-sample chunk at once.  Pairs p = r + 8i
                // (samples 2p, 2p+1), block b = i >> 1.  The 2^-16 input scaling (audio_driver.c:2680-2685)
                // is exact, so it is folded into the correction factors; the Fs/4 translation
                // (freq_shift.c:219-262) is a per-lane sign/swap pattern folded into the same factors.
                {
                    float fi[16], fq[16];
                    int lvmax = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int4 v = *reinterpret_cast<const int4 *>(raw + 2 * (r + 8 * i));
                        lvmax = max(lvmax, max(abs(v.x), abs(v.z)));
                        fi[2 * i] = (float)v.x; fq[2 * i] = (float)v.y; fi[2 * i + 1] = (float)v.z; fq[2 * i + 1] = (float)v.w;
                    }
                    lvmax >>= 16;                                                // audio_driver.c:2662-2675
                    ls.clip |= (lvmax > 1024 ? 1 : 0) | (lvmax > 2048 ? 2 : 0) | (lvmax > 4096 ? 4 : 0);
                    const float kS = 0.0000152587890625f;                        // 2^-16
                    float c1b[4], c2b[4];
                    if (iq_auto) {
                        // Moseley & Slump block statistics (:2274-2279) for the four blocks
                        float s1[4], s2[4], s3[4];
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            s1[b] = 0.0f; s2[b] = 0.0f; s3[b] = 0.0f;
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float vi = fi[4 * b + k], vq = fq[4 * b + k];
                                s1[b] += __fmul_rn(sign_new(vi), vq); s2[b] += fabsf(vi); s3[b] += fabsf(vq);
                            }
                        }
#pragma unroll
                        for (int dlt = 1; dlt < 8; dlt <<= 1) {
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                s1[b] += __shfl_xor_sync(gmask, s1[b], dlt, 8);
                                s2[b] += __shfl_xor_sync(gmask, s2[b], dlt, 8);
                                s3[b] += __shfl_xor_sync(gmask, s3[b], dlt, 8);
                            }
                        }
                        // first-order low-pass over blocks (:2281-2283), then M_c1 / M_c2 (:2285-2295):
                        // lane r computes the pair of block r & 3, the group shares them by shuffle
                        float t1 = ls.te1, t2 = ls.te2, t3 = ls.te3, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
                        const float kE = 0.003f * 0.03125f * kS;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            t1 = fmaf(0.997f, t1, -kE * s1[b]); t2 = fmaf(0.997f, t2, kE * s2[b]); t3 = fmaf(0.997f, t3, kE * s3[b]);
                            if ((r & 3) == b) { m1 = t1; m2 = t2; m3 = t3; }
                        }
                        ls.te1 = t1; ls.te2 = t2; ls.te3 = t3;
                        const float den = m2 * m2;
                        const float c1m = (m2 != 0.0f) ? __fdividef(m1, m2) : 0.0f;
                        const float hlp = (den > 0.0f) ? __fdividef(fmaf(m3, m3, -m1 * m1), den) : den;
                        const float c2m = (hlp > 0.0f) ? hlp * rsqrtf(hlp) : 1.0f;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            c1b[b] = __shfl_sync(gmask, c1m, b, 8);
                            c2b[b] = __shfl_sync(gmask, c2m, b, 8);
                        }
                        ls.c1 = c1b[3]; ls.c2 = c2b[3];
                    }
                    // per-lane Fs/4 pattern: sample 2p (e = 0) has phase (2r) & 3 in {0, 2}, sample 2p+1 phase +1
                    const float sg0 = (shift_kind == 1 && (r & 1)) ? -1.0f : 1.0f;
                    const float sg1 = (shift_kind == 1) ? (shift_down ? -sg0 : sg0) : 1.0f;
                    float *pi0 = xi + ((2 * r) & 3) * XP + XH + (r >> 1), *pq0 = xq + ((2 * r) & 3) * XP + XH + (r >> 1);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int b = i >> 1;
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            float vi = fi[2 * i + e], vq = fq[2 * i + e];
                            if (iq_auto) {
                                vq = fmaf(c1b[b], vi, vq);            // q += M_c1 * i  (:2308-2311)
                                vi = vi * c2b[b];                     // i *= M_c2      (:2313)
                            } else {
                                vi = vi * adj_i; vq = vq * adj_q;     // manual gain / phase (:2259-2267)
                                if (phase_bal < 0.0f) vq = fmaf(vi, phase_bal, vq);
                                else if (phase_bal > 0.0f) vi = fmaf(vq, phase_bal, vi);
                            }
                            float oi, oq;
                            if (e == 0 || shift_kind != 1) { oi = vi * (kS * sg0); oq = vq * (kS * sg0); }
                            else { oi = vq * (kS * sg1); oq = vi * (-kS * sg1); }
                            pi0[e * XP + 4 * i] = oi;
                            pq0[e * XP + 4 * i] = oq;
                        }
                    }
                }
#endif
                __syncwarp();
                // ---- the staging buffer is consumed: fetch the next chunk behind the FIR work ----
                if (t + 1 < nchunks && n_warp_ch > 0) {
                    fence_proxy_async();
                    if (lane == 0) mbar_expect_tx(bar, (unsigned)(n_warp_ch * CH4 * sizeof(int2)));
                    __syncwarp();
                    if (r == 0 && active) bulk_g2s(sm.raw + g * CH4, src + (size_t)(t + 1) * CH4, CH4 * sizeof(int2), bar);
                }
                // ---- decimate: outputs 4r .. 4r+3 for I and Q ----
                {
                    float ai[4], aq[4];
                    decimate4(xi, xq, 4 * r, fc, ai, aq);
                    *reinterpret_cast<float4 *>(di + 200 + 4 * r) = make_float4(ai[0], ai[1], ai[2], ai[3]);
                    *reinterpret_cast<float4 *>(dq + 200 + 4 * r) = make_float4(aq[0], aq[1], aq[2], aq[3]);
                }
                __syncwarp();
                // keep the newest 24 entries of every phase: idx 32..55 -> 0..23 (6 float4 per phase)
                {
                    float4 ki[3], kq[3];
#pragma unroll
                    for (int u = 0; u < 3; u++) {
                        const int e = r + 8 * u;             // 0..23 -> (phase, float4)
                        ki[u] = lds128(xi + (e / 6) * XP + 32 + 4 * (e % 6)); kq[u] = lds128(xq + (e / 6) * XP + 32 + 4 * (e % 6));
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 3; u++) {
                        const int e = r + 8 * u;
                        *reinterpret_cast<float4 *>(xi + (e / 6) * XP + 4 * (e % 6)) = ki[u];
                        *reinterpret_cast<float4 *>(xq + (e / 6) * XP + 4 * (e % 6)) = kq[u];
                    }
                }
                // ---- Hilbert pair + sideband combine ----
                {
                    float hi[4], hq[4];
                    hilbert4(di, dq, 4 * r, fc, hi, hq);
                    float *aud = sm.aud[t & 1];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        aud[(4 * r + j) * SMS + g] = lsb ? __fsub_rn(hi[j], hq[j]) : __fadd_rn(hi[j], hq[j]);
                }
                __syncwarp();
                // slide the Hilbert input: d[0..200) = d[32..232)
                {
                    float4 ki[7], kq[7];
#pragma unroll
                    for (int u = 0; u < 7; u++) {
                        const int e = r + 8 * u;             // float4 index 0..49
                        if (e < 50) { ki[u] = lds128(di + 32 + 4 * e); kq[u] = lds128(dq + 32 + 4 * e); }
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 7; u++) {
                        const int e = r + 8 * u;
                        if (e < 50) { *reinterpret_cast<float4 *>(di + 4 * e) = ki[u]; *reinterpret_cast<float4 *>(dq + 4 * e) = kq[u]; }
                    }
                }
            }
            // ---- write out chunk t-5 (output stage, audio_driver.c:2845-2941) ----
            if (t >= PIPE_DEPTH && active) {
                const int c = t - PIPE_DEPTH;
                const float *o = sm.out[c % 3];
#pragma unroll 4
                for (int i = 0; i < 8; i++) {
                    const int pr = r + 8 * i;
                    const int n = 2 * pr;
                    const bool muted = mute && mute[c * 4 + (i >> 1)];
                    const float v0 = muted ? 0.0f : o[n * SMS + g], v1 = muted ? 0.0f : o[(n + 1) * SMS + g];
                    const int w0 = muted ? 0 : format_audio_word(v0), w1 = muted ? 0 : format_audio_word(v1);
                    dst[(size_t)c * 64 + pr] = make_int4(w0, w0, w1, w1);
                    if (dst_f) dst_f[(size_t)c * 64 + pr] = make_float2(v0, v1);
                }
            }
            __syncthreads();
        }
        // ---- store state ----
        if (active) {
            for (int b = r; b < 4 * XH; b += 8) {
                st->s1_hist_i[b] = xi[(b & 3) * XP + (b >> 2)];
                st->s1_hist_q[b] = xq[(b & 3) * XP + (b >> 2)];
            }
            for (int i = r; i < 200; i += 8) { st->s2_hist_i[i] = di[i]; st->s2_hist_q[i] = dq[i]; }
            int clip = ls.clip;
            clip |= __shfl_xor_sync(gmask, clip, 1, 8); clip |= __shfl_xor_sync(gmask, clip, 2, 8); clip |= __shfl_xor_sync(gmask, clip, 4, 8);
            if (r == 0) {
                st->teta1_old = ls.te1; st->teta2_old = ls.te2; st->teta3_old = ls.te3; st->M_c1 = ls.c1; st->M_c2 = ls.c2;
                if (clip & 1) st->adc_quarter_clip = 1;
                if (clip & 2) st->adc_half_clip = 1;
                if (clip & 4) st->adc_clip = 1;
                st->blocks += a.nblocks;
                if (shift_kind != 0 && st->conversion_freq != p.shift_freq) { st->conversion_freq = p.shift_freq; st->osc_vect_i = 0.0f; st->osc_vect_q = 1.0f; }
            }
        }
        return;
    }

    // ======================= serial warps: one channel per lane =================================
    const int g = lane;
    const bool active = g < n_here;
    const int ch = a.chan_list[cta_first + (active ? g : 0)];
    const ChanParams &p = a.params[ch];
    ChanState *st = a.state + ch;

    if (warp == W_LAT) {
        // ---- 10-stage lattice pre-filter (arm_iir_lattice_f32.c:348-440), front-padded ----
        float k[10], v[11], s[10];
        const int n = p.pre.n, pad = 10 - n;
#pragma unroll
        for (int j = 0; j < 10; j++) {
            k[j] = (j >= pad) ? __ldg(pool + p.pre.k_off + (j - pad)) : 0.0f;
            v[j] = (j >= pad) ? __ldg(pool + p.pre.v_off + (j - pad)) : 0.0f;
            s[j] = (j >= pad) ? st->pre_s[j - pad] : 0.0f;
        }
        v[10] = (n > 0) ? __ldg(pool + p.pre.v_off + n) : 1.0f;
        for (int t = 0; t < nchunks + PIPE_DEPTH; t++) {
            const int c = t - 1;
            if (c >= 0 && c < nchunks) {
                const float *in = sm.aud[c & 1];
                float *out = sm.lat[c & 1];
#pragma unroll 4
                for (int i = 0; i < ND; i++) {
                    float f = in[i * SMS + g], acc = 0.0f, fn = f;
#pragma unroll
                    for (int j = 0; j < 10; j++) {
                        const float gg = s[j];
                        fn = __fsub_rn(f, __fmul_rn(k[j], gg));
                        const float gn = __fadd_rn(__fmul_rn(fn, k[j]), gg);
                        acc = __fadd_rn(acc, __fmul_rn(gn, v[j]));
                        if (j > 0) s[j - 1] = gn;
                        f = fn;
                    }
                    acc = __fadd_rn(acc, __fmul_rn(fn, v[10]));
                    s[9] = fn;
                    out[i * SMS + g] = acc;
                }
            }
            __syncthreads();
        }
        if (active) {
#pragma unroll
            for (int j = 0; j < 10; j++) if (j >= pad) st->pre_s[j - pad] = s[j];
        }
        return;
    }

    if (warp == W_AGC) {
        // ---- WDSP AGC (audio_agc.c:349-595), mono, 12 ksps: 49-sample look-ahead ----
        const AgcP ap = p.agc;
        AgcRun ar = { 0, 0, st->agc_ring_max, st->agc_volts, st->agc_save_volts, st->agc_fast_backaverage,
                      st->agc_hang_backaverage, st->agc_hang_counter, st->agc_decay_type, st->agc_state,
                      st->agc_action, st->agc_hang_action };
        // compact ring: slot (wp - k) & 63 holds x[t-k]; history x[-49..-1] from the 192-slot state ring
        const int in_index = st->agc_in_index;
        for (int kk = 1; kk <= AGC_W; kk++) {
            int idx = in_index - (kk - 1);
            idx %= AGC_RB; if (idx < 0) idx += AGC_RB;
            sm.ring[((RING - kk) & (RING - 1)) * SMS + g] = st->agc_ring[idx];
        }
        // Sliding maximum of |x| over the newest 49 samples (audio_agc.c:409-429 keeps it by rescanning
        // the ring whenever the departing sample was the maximum -- data dependent and divergent).  The
        // same exact maximum comes from the van Herk / Gil-Werman decomposition with blocks = chunks of
        // 32: for the sample at offset o of the current chunk
        //   ring_max = max( prefix_max(cur)[o],  o < 16 ? max(M(prev), suffix(prev2)[16 + o]) : suffix(prev)[o - 16] )
        // smax[s1] / smax[s2]: suffix maxima of the previous / second previous chunk.  History x[-48..-1]
        // fills prev completely and prev2 from offset 16 on (the only part ever read).
        int s1 = 0, s2 = 1;
        {
            float m = 0.0f;
            for (int o = ND - 1; o >= 0; o--) {        // prev: x[-32 + o]
                m = fmaxf(m, fabsf(sm.ring[((RING - (ND - o)) & (RING - 1)) * SMS + g]));
                sm.smax[s1][o * SMS + g] = m;
            }
            m = 0.0f;
            for (int o = ND - 1; o >= 16; o--) {       // prev2: x[-64 + o], only x[-48..-33] exist in the window
                m = fmaxf(m, fabsf(sm.ring[((RING - (2 * ND - o)) & (RING - 1)) * SMS + g]));
                sm.smax[s2][o * SMS + g] = m;
            }
        }
        int wp = 0;          // slot of the sample being written (uniform across lanes)
        // a channel can only be in the hang states (2, 4) or carry decay_type / hang_counter when hang was enabled
        const bool any_hang = __any_sync(0xffffffffu, active && (ap.hang_enable || ar.state == 2 || ar.state == 4 || ar.decay_type != 0 || ar.hang_counter > 0));
        for (int t = 0; t < nchunks + PIPE_DEPTH; t++) {
            const int c = t - 2;
            if (c >= 0 && c < nchunks) {
                const float *in = sm.lat[c & 1];
                float *out = sm.agc[c & 1];
                if (ap.mode == 5) {
                    for (int i = 0; i < ND; i++) out[i * SMS + g] = __fmul_rn(in[i * SMS + g], ap.fixed_gain);   // AGC off (audio_agc.c:354-365)
                } else {
                    const float *S1 = sm.smax[s1], *S2 = sm.smax[s2];
                    const float mprev = S1[g];
                    float pmax = 0.0f;       // prefix maximum inside the current chunk
#pragma unroll 1
                    for (int k8 = 0; k8 < ND; k8 += 8) {
                        // ---- operands of 8 samples: all loads before any store of the group ----
                        float x[8], dly[8], cmx[8], vv[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const int o = k8 + j;
                            x[j] = in[o * SMS + g];
                            dly[j] = sm.ring[((wp + j - AGC_W) & (RING - 1)) * SMS + g];
                            cmx[j] = (o < 16) ? fmaxf(mprev, S2[(16 + o) * SMS + g]) : S1[(o - 16) * SMS + g];
                        }
                        // ---- pass 1 (sample-serial): detector state -> volts ----
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float abs_out = fabsf(dly[j]), abs_in = fabsf(x[j]);
                            pmax = fmaxf(pmax, abs_in);
                            ar.fast_backaverage = __fadd_rn(__fmul_rn(ap.fast_backmult, abs_out), __fmul_rn(ap.onemfast_backmult, ar.fast_backaverage));
                            ar.hang_backaverage = __fadd_rn(__fmul_rn(ap.hang_backmult, abs_out), __fmul_rn(ap.onemhang_backmult, ar.hang_backaverage));
                            ar.ring_max = fmaxf(pmax, cmx[j]);
                            if (ar.hang_counter > 0) --ar.hang_counter;
                            const float dv = __fsub_rn(ar.ring_max, ar.volts);
                            const bool attack = ar.ring_max >= ar.volts;
                            float mult_sel = ap.attack_mult;
                            bool upd = true;
                            int nstate = ar.state;
                            if (!any_hang) {
                                // hang AGC disabled on every channel of this warp (the default,
                                // ui_configuration.c:81): only states 0 / 1 / 3 occur, decay_type stays 0 and
                                // the hang counter stays 0 -- the 5-state machine reduces to selects
                                const bool fast = (ar.state == 0) ? (ar.volts > __fmul_rn(ap.pop_ratio, ar.fast_backaverage))
                                                                  : ((ar.state == 1) && (ar.volts > ar.save_volts));
                                if (attack && ar.state >= 2) ar.save_volts = ar.volts;
                                mult_sel = attack ? ap.attack_mult : (fast ? ap.fast_decay_mult : ap.decay_mult);
                                nstate = attack ? 0 : (fast ? 1 : 3);
                            } else if (attack) {
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
                            ar.state = nstate;
                            if (upd) ar.volts = __fadd_rn(ar.volts, __fmul_rn(dv, mult_sel));
                            if (ar.volts < ap.min_volts) { ar.volts = ap.min_volts; ar.action = 0; } else { ar.action = 1; }
                            vv[j] = ar.volts;
                        }
                        // ---- pass 2 (independent per sample): log-slope gain on the delayed sample (:563-570) ----
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float vo = log10f_fast(__fmul_rn(ap.inv_max_input, vv[j]));
                            if (vo > 0.0f) vo = 0.0f;
#if UHSDR_EXACT
                            const float mult = __fdiv_rn(__fsub_rn(ap.out_target, __fmul_rn(ap.slope_constant, vo)), vv[j]);
#else
                            const float mult = __fdividef(fmaf(-ap.slope_constant, vo, ap.out_target), vv[j]);
#endif
                            out[(k8 + j) * SMS + g] = __fmul_rn(dly[j], mult);
                            sm.ring[((wp + j) & (RING - 1)) * SMS + g] = x[j];
                        }
                        wp = (wp + 8) & (RING - 1);
                    }
                    ar.hang_action = (ar.hang_backaverage > ap.hang_level) ? 1 : 0;
                    // suffix maxima of this chunk replace those of prev2; then the roles rotate
                    {
                        float *Sn = sm.smax[s2];
                        float m = 0.0f;
#pragma unroll 8
                        for (int o = ND - 1; o >= 0; o--) {
                            m = fmaxf(m, fabsf(sm.ring[((wp - (ND - o)) & (RING - 1)) * SMS + g]));
                            Sn[o * SMS + g] = m;
                        }
                        const int tmp = s1; s1 = s2; s2 = tmp;
                    }
                }
            }
            __syncthreads();
        }
        if (active && ap.mode != 5) {
            const long long T = (long long)nchunks * ND;
            int new_in = (int)(((long long)st->agc_in_index + T) % AGC_RB);
            int new_out = (int)((((long long)st->agc_out_index + T) % AGC_RB + AGC_RB) % AGC_RB);
            for (int kk = 1; kk <= AGC_W; kk++) {
                int idx = new_in - (kk - 1);
                idx %= AGC_RB; if (idx < 0) idx += AGC_RB;
                st->agc_ring[idx] = sm.ring[((wp - kk) & (RING - 1)) * SMS + g];
            }
            st->agc_in_index = new_in; st->agc_out_index = new_out;
            st->agc_ring_max = ar.ring_max; st->agc_volts = ar.volts; st->agc_save_volts = ar.save_volts;
            st->agc_fast_backaverage = ar.fast_backaverage; st->agc_hang_backaverage = ar.hang_backaverage;
            st->agc_hang_counter = ar.hang_counter; st->agc_decay_type = ar.decay_type; st->agc_state = ar.state;
            st->agc_action = ar.action; st->agc_hang_action = ar.hang_action;
        }
        return;
    }

    if (warp == W_EQ) {
        // ---- fixed gain (:2513-2524), biquad_1 (:2527), x4 interpolator (:2560-2577) ----
        float bc[4][5]; BiquadS bs[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
#pragma unroll
            for (int q = 0; q < 5; q++) bc[s][q] = p.bq1[s][q];
            bs[s] = st->bq1[s];
        }
        const float scale_gain = p.scale_gain;
        // interpolator taps as [phase j][k] with phase length padded to 4 (leading zeros)
        float ic[4][4], ih[3];
        const int P = p.interp_plen;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int kq = 0; kq < 4; kq++) {
                const int kk = kq - (4 - P);
                ic[j][kq] = (kk >= 0) ? __ldg(pool + p.interp_c + (3 - j) + 4 * kk) : 0.0f;
            }
#pragma unroll
        for (int q = 0; q < 3; q++) ih[q] = st->interp_hist[INTERP_HIST - 3 + q];
        for (int t = 0; t < nchunks + PIPE_DEPTH; t++) {
            const int c = t - 3;
            if (c >= 0 && c < nchunks) {
                const float *in = sm.agc[c & 1];
                float *out = sm.out[c % 3];
#pragma unroll 2
                for (int i = 0; i < ND; i++) {
                    float x = __fmul_rn(in[i * SMS + g], scale_gain);
#pragma unroll
                    for (int s = 0; s < 4; s++) x = biquad_step(x, bc[s], bs[s]);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float sum = 0.0f;
                        sum = mad(ih[0], ic[j][0], sum); sum = mad(ih[1], ic[j][1], sum);
                        sum = mad(ih[2], ic[j][2], sum); sum = mad(x, ic[j][3], sum);
                        out[(4 * i + j) * SMS + g] = sum;
                    }
                    ih[0] = ih[1]; ih[1] = ih[2]; ih[2] = x;
                }
            }
            __syncthreads();
        }
        if (active) {
#pragma unroll
            for (int s = 0; s < 4; s++) st->bq1[s] = bs[s];
            // the generic kernel keeps the newest 8 decimated samples; older slots are only read
            // by interpolators with longer phases than this kernel is eligible for
            for (int q = 0; q < INTERP_HIST - 3; q++) st->interp_hist[q] = 0.0f;
#pragma unroll
            for (int q = 0; q < 3; q++) st->interp_hist[INTERP_HIST - 3 + q] = ih[q];
        }
        return;
    }

    // warp == W_POST
    {
        // ---- anti-alias lattice (:2581-2583, 6 stages, some paths), treble biquad (:2832), x10 ----
        float k[6], v[7], s[6];
        const int n = p.aa.n;
#pragma unroll
        for (int j = 0; j < 6; j++) {
            k[j] = (n == 6) ? __ldg(pool + p.aa.k_off + j) : 0.0f;
            v[j] = (n == 6) ? __ldg(pool + p.aa.v_off + j) : 0.0f;
            s[j] = (n == 6) ? st->aa_s[j] : 0.0f;
        }
        v[6] = (n == 6) ? __ldg(pool + p.aa.v_off + 6) : 1.0f;
        const bool any_aa = __any_sync(0xffffffffu, active && n == 6);
        float bc[5];
#pragma unroll
        for (int q = 0; q < 5; q++) bc[q] = p.bq2[q];
        BiquadS bs = st->bq2;
        for (int t = 0; t < nchunks + PIPE_DEPTH; t++) {
            const int c = t - 4;
            if (c >= 0 && c < nchunks) {
                float *buf = sm.out[c % 3];
#pragma unroll 4
                for (int i = 0; i < CH4; i++) {
                    float x = buf[i * SMS + g];
                    if (any_aa) {
                        float f = x, acc = 0.0f, fn = x;
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            const float gg = s[j];
                            fn = __fsub_rn(f, __fmul_rn(k[j], gg));
                            const float gn = __fadd_rn(__fmul_rn(fn, k[j]), gg);
                            acc = __fadd_rn(acc, __fmul_rn(gn, v[j]));
                            if (j > 0) s[j - 1] = gn;
                            f = fn;
                        }
                        acc = __fadd_rn(acc, __fmul_rn(fn, v[6]));
                        s[5] = fn;
                        x = (n == 6) ? acc : x;
                    }
                    x = biquad_step(x, bc, bs);
                    buf[i * SMS + g] = __fmul_rn(x, 10.0f);       // LINE_OUT_SCALING_FACTOR (:2860)
                }
            }
            __syncthreads();
        }
        if (active) {
            if (n == 6) {
#pragma unroll
                for (int j = 0; j < 6; j++) st->aa_s[j] = s[j];
            }
            st->bq2 = bs;
        }
    }
}

// A channel runs on the fused kernel when its chain is the narrow-SSB/CW topology with the stock
// 83-tap decimator and 199-tap Hilbert pair, Fs/4 (or no) translation, a x4 interpolator with at
// most 4 taps per phase, and no deferred consumer (spectral NR, spectrum ring).
bool fused_eligible(const ChanParams &p)
{
    return p.configured && p.topo == TOPO_SSB_DEC_FIRST && p.M == 4 && p.s1_ntaps == 83 && p.s2_ntaps == 199 &&
           p.shift_kind != 2 && !p.nr_enable && !p.notch_enable && !p.spectrum_enable && p.pre.n <= 10 && (p.aa.n == 0 || p.aa.n == 6) &&
           p.interp_L == 4 && p.interp_plen >= 1 && p.interp_plen <= 4 && p.agc.attack_buffsize == AGC_W;
}

bool fused_eligible_ext(const ChanParams &p)
{
    return p.configured && p.topo == TOPO_SSB_DEC_FIRST && p.M == 4 && p.s1_ntaps == 83 && p.s2_ntaps == 199 &&
           p.shift_kind != 2 && !p.notch_enable && (!p.spectrum_enable || p.zoom_m == 0) && p.pre.n <= 10 && (p.aa.n == 0 || p.aa.n == 6) &&
           p.interp_L == 4 && p.interp_plen >= 1 && p.interp_plen <= 4 && p.agc.attack_buffsize == AGC_W &&
           (!p.nr_enable || rx_serial2_eligible(p)) &&
           !p.nb_enable;     // the LPC noise blanker is a chain of threshold decisions: it keeps the FP32 FIRs of the split path in front of it
}

void fill_fused_coefs(FusedCoefs *fc, const float *dec83, const float *hil_i199, const float *hil_q199)
{
    memset(fc, 0, sizeof(*fc));
    memcpy(fc->dec + DEC_PAD, dec83, 83 * sizeof(float));
    memcpy(fc->hil_i + HIL_PAD, hil_i199, 199 * sizeof(float));
    memcpy(fc->hil_q + HIL_PAD, hil_q199, 199 * sizeof(float));
}

cudaError_t launch_rx_ssb_fused(const RxArgs &a, const FusedCoefs &fc, int sm_count, cudaStream_t stream)
{
    if (a.num_items <= 0) return cudaSuccess;
    if (a.nblocks % 4 != 0 || a.chan_list == nullptr) return cudaErrorInvalidValue;
    // channels per CTA: fill every SM once, in multiples of 4
    int per = (a.num_items + sm_count - 1) / sm_count;
    per = ((per + 3) / 4) * 4;
    if (per > FG) per = FG;
    if (per < 4) per = 4;
    const int grid = (a.num_items + per - 1) / per;
    cudaError_t e = cudaFuncSetAttribute(rx_ssb_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    rx_ssb_fused_kernel<<<grid, NTHREADS, sizeof(Smem), stream>>>(a, fc, per);
    return cudaGetLastError();
}

}  // namespace uhsdr
