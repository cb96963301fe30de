// rx_ssb_tc.cu -- fused narrow-SSB/CW receiver kernel with the 199-tap Hilbert pair on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).  Shipping build only; the exact
// build keeps the CUDA-core kernel of rx_ssb_fused.cu.
//
// Chain (FilterPathInfo[4..47], mchf-eclipse/drivers/audio/audio_filter.c:147-922; flow in
// audio_driver.c:2603-2942):
//   format + IQ correction + Fs/4 translate -> 83-tap /4 decimator on I and Q      (FIR warps, FP32 pipe)
//   -> 199-tap Hilbert pair @12 ksps -> I +/- Q                                    (tensor cores)
//   -> 10-stage lattice IIR | WDSP AGC | gain, 4-stage biquad | x4 interpolator, (anti-alias lattice),
//      treble biquad, x10, int32 << 16                                             (four serial warps)
//
// One persistent CTA owns up to 28 channels for the whole launch; every sample crosses HBM once.
//
// Hilbert pair as a Toeplitz GEMM (north_star form (1)).  arm_fir_f32 (arm_fir_f32.c:522-529) computes
// y[n] = sum_k c[k] d[n - 198 + k].  For a chunk of 64 outputs and all channels of the CTA
//   D[m][ch] = sum_p T[m][p] X[p][ch],   T[m][p] = c[p - m - 10],   p = 0..271 (window of 272 inputs)
// is a 64 x 32 x 272 GEMM whose A operand is the same for every chunk and channel.  Because T is
// Toeplitz, the 64 x 16 slab of k-step kk is rows [264 - 16 kk, +64) of ONE table
// G[r][q] = c[q - r + 254] (328 rows x 16), so the A descriptor just slides through a 10.5 KB table.
// The B operand is the decimator output itself: a channel-major ring in shared memory in the
// canonical K-major (no-swizzle) layout, where advancing in time is again an address offset.
// I and Q accumulate into the same D (USB: I + Q; LSB: the Q samples are stored negated).
// Arithmetic: BF16 operands, FP32 accumulation, split x = x1 + x2 (16 significant bits), c = c1 + c2,
// D = c1 x1 + c2 x1 + c1 x2  -- relative error ~4e-6 (106 dB), measured in scripts/micro/umma_toeplitz.cu.
#include <cuda_bf16.h>

#include <type_traits>

#include "dsp_device.cuh"
#include "kernels.h"
#include "uhsdr_b200.h"

#if !UHSDR_EXACT

namespace uhsdr {

namespace {

constexpr int FG = 28;             // channel slots per CTA (7 FIR warps x 4 channels)
constexpr int CH4 = 128;           // input samples per step (4 blocks)
constexpr int ND = 32;             // decimated samples per step
constexpr int XP = 56;             // per-phase slots of the decimator staging: 24 history (21 used) + 32 new
constexpr int XH = 24;
constexpr int XCH = 2 * 4 * XP + 4;
constexpr int SMS = 29;            // channel-minor stride of the serial-stage queues (odd: conflict free)
constexpr int AGC_W = 49;          // attack_buffsize at 12 ksps (audio_agc.c:290)
constexpr int LR = 5 * ND;          // rows of the lattice-output ring (decimated samples)
constexpr int AG = 8;              // AGC samples per group (operands loaded together, detector serial, gain law parallel)
constexpr int NWARP_FIR = FG / 4;
constexpr int DEC_PAD = 32;        // FusedCoefs::dec carries the 83 taps at [32, 115)
// tensor-core Hilbert
constexpr int HT = 336;            // ring time slots per channel (272-slot window + 64 being produced)
constexpr int HWIN = 272;          // window of one 64-output chunk (17 k-steps of 16)
constexpr int KSTEPS = HWIN / 16;
constexpr int GROWS = 328;         // Toeplitz table rows
constexpr int RING_BYTES = 4 * (HT / 8) * 128;     // one array: [channel group 0..3][time/8][channel%8][time%8] bf16
constexpr int G_BYTES = GROWS * 32;                // [row/8][k half][row%8][8] bf16
constexpr int TMEM_COLS = 64;      // two 64 x 32 fp32 accumulators
// warp roles
// (warp id % 4 is the scheduler: the four serial warps and the MMA issuer are spread over all four)
constexpr int W_AGC = NWARP_FIR, W_POST = NWARP_FIR + 1, W_LAT = NWARP_FIR + 2, W_MMA = NWARP_FIR + 3, W_BQ = NWARP_FIR + 4;
constexpr int NTHREADS = 32 * (NWARP_FIR + 5);
// software pipeline, in steps of 128 input samples: step h is decimated at iteration h, its Hilbert
// outputs leave TMEM at h + 4, lattice h + 5, AGC h + 6, gain + biquad cascade h + 7,
// interpolator / treble / output formatting h + 8
constexpr int IT_EPI = 4, IT_LAT = 5, IT_AGC = 6, IT_BQ = 7, IT_POST = 8;
constexpr int PIPE_DEPTH = IT_POST;

struct Smem {
    alignas(128) unsigned char ring[4][RING_BYTES];   // I1, I2, Q1, Q2: bf16 split of the decimator outputs
    alignas(128) unsigned char g[4][G_BYTES];         // Toeplitz tables: hil_i c1, c2, hil_q c1, c2
    alignas(16) float x[FG * XCH];                    // decimator staging, polyphase layout
    float aud[2][ND * SMS];
    float lat[LR * SMS];              // lattice output, a ring of 5 steps: AGC detector and gain stage read x[n-49] from it
    float agc[2][ND * SMS];           // AGC "volts" per sample (detector -> gain stage)
    float bq[2][ND * SMS];
    float smax[2][ND * SMS];
    alignas(8) unsigned long long mma_bar[2];
    unsigned tmem_base;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }

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

// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B; lbo = byte distance
// between the two 16-byte K halves of a k-step, sbo = byte distance between 8-row groups
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr, unsigned lbo, unsigned sbo)
{
    return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// x rounded to 16 significant bits (round to nearest even) and split into two bf16 values whose sum
// is that rounded value exactly: hi = upper 16 bits (truncation), lo = remainder (<= 8 significant bits).
// The rounded value is what the channel state keeps, so the split is reproducible across calls.
__device__ __forceinline__ float round16(float x)
{
    unsigned u = __float_as_uint(x);
    u += 0x7fu + ((u >> 8) & 1u);
    return __uint_as_float(u & 0xffffff00u);
}
__device__ __forceinline__ void split_bf16(float h, unsigned &hi, unsigned &lo)
{
    const unsigned u = __float_as_uint(h);
    hi = u >> 16;
    lo = __float_as_uint(h - __uint_as_float(u & 0xffff0000u)) >> 16;
}

// byte offset of (channel slot g, time slot s) inside one ring array
__device__ __forceinline__ int ring_off(int g, int s) { return (g >> 3) * ((HT / 8) * 128) + (s >> 3) * 128 + (g & 7) * 16 + (s & 7) * 2; }

struct FirLaneState {
    float te1, te2, te3;     // teta*_old
    float c1, c2;            // M_c1, M_c2
    int clip;                // bit0 quarter, bit1 half, bit2 full
};

template <int I, int N, typename F> __device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}

// Decimator: y[m] = sum_k c[k] x[4m - 82 + k] (arm_fir_decimate_f32.c:455-486).  With 96 history slots
// in front (24 per phase), buffer position b = 4m + k + 14; phase = b & 3, idx = b >> 2.  Each lane
// makes 4 consecutive outputs m0..m0+3 of one signal (the caller loops over I and Q with the same code, which
// halves the instruction-cache footprint); element (q, e, ph) of the 7 x 4 float4 loads
// is position 4 (m0 + 4q + e) + ph, i.e. tap K = 16q + 4 (e - j) + ph - 14 of output j.  Fully unrolled:
// every tap is an immediate constant-bank operand and taps outside [0, 82] generate no instruction.
__device__ __forceinline__ void decimate4(const float *xp, int m0, const FusedCoefs &fc, float acc[4])
{
#pragma unroll
    for (int j = 0; j < 4; j++) acc[j] = 0.0f;
    float4 v[2][4];                       // the loads of group q + 1 are issued before the FMAs of group q
#pragma unroll
    for (int ph = 0; ph < 4; ph++) v[0][ph] = lds128(xp + ph * XP + m0);
    static_for<0, 7>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        if constexpr (q < 6) {
#pragma unroll
            for (int ph = 0; ph < 4; ph++) v[(q + 1) & 1][ph] = lds128(xp + ph * XP + m0 + 4 * (q + 1));
        }
        static_for<0, 4>([&](auto ec) {
            constexpr int e = decltype(ec)::value;
            static_for<0, 4>([&](auto pc) {
                constexpr int ph = decltype(pc)::value;
                const float4 vv = v[q & 1][ph];
                const float x = (e == 0) ? vv.x : (e == 1) ? vv.y : (e == 2) ? vv.z : vv.w;
                static_for<0, 4>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    constexpr int K = 16 * q + 4 * (e - j) + ph - 14;
                    if constexpr (K >= 0 && K < 83) acc[j] = fmaf(fc.dec[DEC_PAD + K], x, acc[j]);
                });
            });
        });
    });
}

}  // namespace

__global__ void __launch_bounds__(NTHREADS, 1)
rx_ssb_tc_kernel(const __grid_constant__ RxArgs a, const __grid_constant__ FusedCoefs fc, int chans_per_cta, int hil_ci, int hil_cq)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cta_first = blockIdx.x * chans_per_cta;
    const int n_here = min(chans_per_cta, a.num_items - cta_first);
    const int nsteps = a.nblocks / 4;
    const int niter = nsteps + PIPE_DEPTH;
    const float *__restrict__ pool = a.pool;

    // ---- one-time setup by all threads: Toeplitz tables, zeroed ring, barriers, TMEM ----
    for (int i = threadIdx.x; i < 2 * GROWS * 16; i += NTHREADS) {
        const int which = i / (GROWS * 16), e = i % (GROWS * 16);
        const int r = e >> 4, q = e & 15;
        const int kidx = q - r + 254;
        const float cv = (kidx >= 0 && kidx < 199) ? __ldg(pool + (which ? hil_cq : hil_ci) + kidx) : 0.0f;
        const __nv_bfloat16 c1 = __float2bfloat16_rn(cv);
        const __nv_bfloat16 c2 = __float2bfloat16_rn(cv - __bfloat162float(c1));
        const int off = (r >> 3) * 256 + (q >> 3) * 128 + (r & 7) * 16 + (q & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(sm.g[2 * which] + off) = c1;
        *reinterpret_cast<__nv_bfloat16 *>(sm.g[2 * which + 1] + off) = c2;
    }
    for (int i = threadIdx.x; i < 4 * RING_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm.ring)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.mma_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.mma_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // tables and zeroed ring -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;

    if (warp < NWARP_FIR) {
        // ======================= FIR warp: 4 channels x 8 lanes ==================================
        const int cl = lane >> 3, r = lane & 7;
        const int g = warp * 4 + cl;                   // channel slot in the CTA
        const bool active = g < n_here;
        const int ch = active ? a.chan_list[cta_first + g] : a.chan_list[cta_first];
        const ChanParams &p = a.params[ch];
        ChanState *st = a.state + ch;
        float *xi = sm.x + g * XCH, *xq = xi + 4 * XP;
        const unsigned gmask = 0xffu << (8 * cl);
        const int lsb = p.lsb;

        FirLaneState ls;
        ls.te1 = st->teta1_old; ls.te2 = st->teta2_old; ls.te3 = st->teta3_old; ls.c1 = st->M_c1; ls.c2 = st->M_c2; ls.clip = 0;
        for (int b = r; b < 4 * XH; b += 8) {
            xi[(b & 3) * XP + (b >> 2)] = active ? st->s1_hist_i[b] : 0.0f;
            xq[(b & 3) * XP + (b >> 2)] = active ? st->s1_hist_q[b] : 0.0f;
        }
        // Hilbert history d[-198..-1] (state slots 2..199) -> ring slots 10..207
        if (active) {
            for (int i = 2 + r; i < 200; i += 8) {
                const float hi_ = round16(st->s2_hist_i[i]);
                const float hq_ = round16(lsb ? -st->s2_hist_q[i] : st->s2_hist_q[i]);
                unsigned i1, i2, q1, q2;
                split_bf16(hi_, i1, i2); split_bf16(hq_, q1, q2);
                const int off = ring_off(g, i + 8);
                *reinterpret_cast<unsigned short *>(sm.ring[0] + off) = (unsigned short)i1;
                *reinterpret_cast<unsigned short *>(sm.ring[1] + off) = (unsigned short)i2;
                *reinterpret_cast<unsigned short *>(sm.ring[2] + off) = (unsigned short)q1;
                *reinterpret_cast<unsigned short *>(sm.ring[3] + off) = (unsigned short)q2;
            }
        }
        const int iq_auto = p.iq_auto, shift_kind = p.shift_kind, shift_down = p.shift_down;
        const bool any_auto = __any_sync(0xffffffffu, iq_auto != 0);
        const bool fast_fe = __all_sync(0xffffffffu, iq_auto != 0 && shift_kind == 1);
        const float adj_i = p.adj_i, adj_q = p.adj_q, phase_bal = p.phase_bal;
        const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
        const int4 *__restrict__ src = reinterpret_cast<const int4 *>(reinterpret_cast<const int2 *>(a.iq) + chan_base);

        // input prefetch: 8 x int4 = the 16 consecutive samples 16r .. 16r+15 of the step, one step ahead, straight from global memory
        int4 pre[8];
        if (nsteps > 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) pre[i] = active ? __ldg(src + 8 * r + i) : make_int4(0, 0, 0, 0);
        }
        __syncwarp();

        for (int t = 0; t < niter; t++) {
            // the ring slots written in this step were last read by the MMAs of chunk t/2 - 2
            if (t < nsteps && (t & 1) == 0 && t >= 4) mbar_wait(&sm.mma_bar[((t >> 1) - 2) & 1], (unsigned)((((t >> 1) - 2) >> 1) & 1));
            // ---- Hilbert outputs of step t - 4 leave TMEM: rows 0..31 via quadrants 0,1 (warps 0,1), rows 32..63 via 2,3 ----
            {
                const int h = t - IT_EPI;
                if (warp < 4 && h >= 0 && h < nsteps && (warp >> 1) == (h & 1)) {
                    const int chunk = h >> 1;
                    mbar_wait(&sm.mma_bar[chunk & 1], (unsigned)((chunk >> 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    unsigned v[32];
                    const unsigned taddr = tmem + (unsigned)((chunk & 1) * 32) + ((unsigned)(warp * 32) << 16);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (lane < 16) {
                        float *aud = sm.aud[h & 1] + ((warp & 1) * 16 + lane) * SMS;
#pragma unroll
                        for (int c = 0; c < FG; c++) aud[c] = __uint_as_float(v[c]);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
            }
            if (t < nsteps) {
                // ---- front end: the whole 128-sample step at once.  Lane r owns the 16 consecutive samples 16r .. 16r+15,
                // i.e. one half of block r >> 1, and stores them as one float4 per decimator phase.  The 2^-16 input
                // scaling (audio_driver.c:2680-2685) is exact, so it is folded into the correction factors; the Fs/4
                // translation (freq_shift.c:219-262) is a sign/swap pattern of period 4 folded into the same factors.
                {
                    float fi[16], fq[16];
                    int lvmax = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int4 v = pre[i];
                        lvmax = max(lvmax, max(abs(v.x), abs(v.z)));
                        fi[2 * i] = (float)v.x; fq[2 * i] = (float)v.y; fi[2 * i + 1] = (float)v.z; fq[2 * i + 1] = (float)v.w;
                    }
                    // fetch the next step behind the FIR work
                    if (t + 1 < nsteps && active) {
#pragma unroll
                        for (int i = 0; i < 8; i++) pre[i] = __ldg(src + (size_t)(t + 1) * 64 + 8 * r + i);
                    }
                    lvmax >>= 16;                                                // audio_driver.c:2662-2675
                    ls.clip |= (lvmax > 1024 ? 1 : 0) | (lvmax > 2048 ? 2 : 0) | (lvmax > 4096 ? 4 : 0);
                    const float kS = 0.0000152587890625f;                        // 2^-16
                    float c1m = 0.0f, c2m = 1.0f;                                // M_c1, M_c2 of this lane's block
                    if (any_auto) {
                        // Moseley & Slump block statistics (:2274-2279); sign(i) * q as a sign-bit transfer (differs from
                        // Math_sign_new only for i == 0, where the term is +-q instead of 0)
                        float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            s1 += __uint_as_float(__float_as_uint(fq[k]) ^ (__float_as_uint(fi[k]) & 0x80000000u));
                            s2 += fabsf(fi[k]); s3 += fabsf(fq[k]);
                        }
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 1, 8); s2 += __shfl_xor_sync(0xffffffffu, s2, 1, 8); s3 += __shfl_xor_sync(0xffffffffu, s3, 1, 8);
                        // first-order low-pass over the four blocks (:2281-2283), then M_c1 / M_c2 (:2285-2295) of the own block
                        float t1 = ls.te1, t2 = ls.te2, t3 = ls.te3, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
                        const float kE = 0.003f * 0.03125f * kS;
                        float bs1[4], bs2[4], bs3[4];           // all twelve broadcasts in flight before the recurrence uses them
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            bs1[b] = __shfl_sync(0xffffffffu, s1, 2 * b, 8); bs2[b] = __shfl_sync(0xffffffffu, s2, 2 * b, 8); bs3[b] = __shfl_sync(0xffffffffu, s3, 2 * b, 8);
                        }
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            t1 = fmaf(0.997f, t1, -kE * bs1[b]); t2 = fmaf(0.997f, t2, kE * bs2[b]); t3 = fmaf(0.997f, t3, kE * bs3[b]);
                            if ((r >> 1) == b) { m1 = t1; m2 = t2; m3 = t3; }
                        }
                        const float den = m2 * m2;
                        const float hlp = (den > 0.0f) ? __fdividef(fmaf(m3, m3, -m1 * m1), den) : den;
                        if (iq_auto) {
                            ls.te1 = t1; ls.te2 = t2; ls.te3 = t3;
                            c1m = (m2 != 0.0f) ? __fdividef(m1, m2) : 0.0f;
                            c2m = (hlp > 0.0f) ? hlp * rsqrtf(hlp) : 1.0f;
                            ls.c1 = c1m; ls.c2 = c2m;                            // lanes 6, 7 hold the block-3 values the state keeps
                        }
                    }
                    float oi[16], oq[16];
                    if (fast_fe) {
                        // every channel of the warp: automatic IQ correction + Fs/4 translation (the default).
                        //   i' = c2 i, q' = q + c1 i;  phase 0: (i', q')  1: (q', -i')  2: (-i', -q')  3: (-q', i'), (x sgd when translating down)
                        const float sgd = shift_down ? -1.0f : 1.0f;
                        const float fa = c2m * kS, fd = c1m * kS, fas = fa * sgd, fds = fd * sgd, ks = kS * sgd;
#pragma unroll
                        for (int k = 0; k < 16; k += 4) {
                            oi[k] = fi[k] * fa;                                   oq[k] = fmaf(fi[k], fd, fq[k] * kS);
                            oi[k + 1] = fmaf(fi[k + 1], fds, fq[k + 1] * ks);     oq[k + 1] = fi[k + 1] * -fas;
                            oi[k + 2] = fi[k + 2] * -fa;                          oq[k + 2] = fmaf(fi[k + 2], -fd, fq[k + 2] * -kS);
                            oi[k + 3] = fmaf(fi[k + 3], -fds, fq[k + 3] * -ks);   oq[k + 3] = fi[k + 3] * fas;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            float vi = fi[k], vq = fq[k];
                            if (iq_auto) {
                                vq = fmaf(c1m, vi, vq);               // q += M_c1 * i  (:2308-2311)
                                vi = vi * c2m;                        // i *= M_c2      (:2313)
                            } else {
                                vi = vi * adj_i; vq = vq * adj_q;     // manual gain / phase (:2259-2267)
                                if (phase_bal < 0.0f) vq = fmaf(vi, phase_bal, vq);
                                else if (phase_bal > 0.0f) vi = fmaf(vq, phase_bal, vi);
                            }
                            vi *= kS; vq *= kS;
                            if (shift_kind == 1) {
                                const float sgd = shift_down ? -1.0f : 1.0f;
                                const int ph = k & 3;
                                const float ti = vi, tq_ = vq;
                                if (ph == 1) { vi = tq_ * sgd; vq = -ti * sgd; }
                                else if (ph == 2) { vi = -ti; vq = -tq_; }
                                else if (ph == 3) { vi = -tq_ * sgd; vq = ti * sgd; }
                            }
                            oi[k] = vi; oq[k] = vq;
                        }
                    }
                    // sample 16r + 4k + ph -> phase array ph, slot XH + 4r + k: one float4 per phase
#pragma unroll
                    for (int ph = 0; ph < 4; ph++) {
                        *reinterpret_cast<float4 *>(xi + ph * XP + XH + 4 * r) = make_float4(oi[ph], oi[ph + 4], oi[ph + 8], oi[ph + 12]);
                        *reinterpret_cast<float4 *>(xq + ph * XP + XH + 4 * r) = make_float4(oq[ph], oq[ph + 4], oq[ph + 8], oq[ph + 12]);
                    }
                }
                __syncwarp();
                // ---- decimate: outputs 4r .. 4r+3 for I and Q -> 16-bit rounding, bf16 split, Hilbert ring ----
                {
                    const int sl = (208 + ND * t + 4 * r) % HT;          // ring slot of decimated sample 32 t + 4 r
                    const int off = ring_off(g, sl);
#pragma unroll 1
                    for (int c = 0; c < 2; c++) {
                        float acc[4];
                        decimate4(c ? xq : xi, 4 * r, fc, acc);
                        const bool neg = c && lsb;                         // LSB: I - Q, the Q samples are stored negated
                        unsigned h1[4], h2[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) split_bf16(round16(neg ? -acc[j] : acc[j]), h1[j], h2[j]);
                        *reinterpret_cast<uint2 *>(sm.ring[2 * c] + off) = make_uint2(h1[0] | (h1[1] << 16), h1[2] | (h1[3] << 16));
                        *reinterpret_cast<uint2 *>(sm.ring[2 * c + 1] + off) = make_uint2(h2[0] | (h2[1] << 16), h2[2] | (h2[3] << 16));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // visible to the tensor core after the step barrier
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
            }
            __syncthreads();
        }
        // ---- store state: decimator history, the newest 198 Hilbert inputs, IQ-correction state ----
        if (active) {
            for (int b = r; b < 4 * XH; b += 8) {
                st->s1_hist_i[b] = xi[(b & 3) * XP + (b >> 2)];
                st->s1_hist_q[b] = xq[(b & 3) * XP + (b >> 2)];
            }
            // state slot i (2..199) = decimated sample 32 nsteps - 200 + i = ring slot (208 + 32 nsteps - 200 + i) % HT
            for (int i = 2 + r; i < 200; i += 8) {
                const int off = ring_off(g, (8 + ND * nsteps + i) % HT);
                const unsigned i1 = *reinterpret_cast<const unsigned short *>(sm.ring[0] + off), i2 = *reinterpret_cast<const unsigned short *>(sm.ring[1] + off);
                const unsigned q1 = *reinterpret_cast<const unsigned short *>(sm.ring[2] + off), q2 = *reinterpret_cast<const unsigned short *>(sm.ring[3] + off);
                const float vi = __uint_as_float(i1 << 16) + __uint_as_float(i2 << 16);
                const float vq = __uint_as_float(q1 << 16) + __uint_as_float(q2 << 16);
                st->s2_hist_i[i] = vi;
                st->s2_hist_q[i] = lsb ? -vq : vq;
            }
            int clip = ls.clip;
            ls.c1 = __shfl_sync(gmask, ls.c1, 6, 8); ls.c2 = __shfl_sync(gmask, ls.c2, 6, 8);
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

    if (warp == W_MMA) {
        // ======================= MMA issue: one elected lane ====================================
        // Chunk c (steps 2c, 2c+1) is complete after iteration min(2c+1, nsteps-1); its 102 MMAs are issued at
        // the start of the next iteration into accumulator c & 1 and committed to mma_bar[c & 1].
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(32 >> 3) << 17) | ((unsigned)(64 >> 4) << 24);
        const unsigned sbo_b = (HT / 8) * 128;
        for (int t = 0; t < niter; t++) {
            int chunk = -1;
            if (t >= 1 && t <= nsteps) {
                if ((t & 1) == 0) chunk = (t >> 1) - 1;
                else if (t == nsteps) chunk = (t - 1) >> 1;
            }
            if (chunk >= 0 && lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d_tmem = tmem + (unsigned)((chunk & 1) * 32);
                const int s0 = (64 * chunk) % HT;                   // ring slot of the window start (decimated sample 64c - 208)
                // k-step kk: A = rows [264 - 16 kk, +64) of the Toeplitz table (start address - 512 B per step),
                //            B = ring slots [s0 + 16 kk, +16) (start address + 256 B per step, wrapping at HT slots)
#pragma unroll 1
                for (int arr = 0; arr < 2; arr++) {
                    unsigned long long a1 = umma_desc(smem_u32(sm.g[2 * arr]) + 264u * 32u, 128, 256);
                    unsigned long long a2 = umma_desc(smem_u32(sm.g[2 * arr + 1]) + 264u * 32u, 128, 256);
                    unsigned long long b1 = umma_desc(smem_u32(sm.ring[2 * arr]) + (unsigned)(s0 >> 3) * 128u, 128, sbo_b);
                    unsigned long long b2 = umma_desc(smem_u32(sm.ring[2 * arr + 1]) + (unsigned)(s0 >> 3) * 128u, 128, sbo_b);
                    int s = s0;
#pragma unroll 1
                    for (int kk = 0; kk < KSTEPS; kk++) {
                        umma_bf16(d_tmem, a1, b1, idesc, (arr > 0 || kk > 0) ? 1u : 0u);
                        umma_bf16(d_tmem, a2, b1, idesc, 1u);
                        umma_bf16(d_tmem, a1, b2, idesc, 1u);
                        a1 -= 512u >> 4; a2 -= 512u >> 4;
                        s += 16;
                        if (s >= HT) { s -= HT; b1 -= (unsigned long long)(((HT - 16) / 8) * 128 >> 4); b2 -= (unsigned long long)(((HT - 16) / 8) * 128 >> 4); }
                        else { b1 += 256u >> 4; b2 += 256u >> 4; }
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.mma_bar[chunk & 1])) : "memory");
            }
            __syncwarp();
            __syncthreads();
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
        return;
    }

    // ======================= serial warps: one channel per lane =================================
    const int g = lane;
    const bool active = g < n_here;
    const int ch = a.chan_list[cta_first + (active ? g : 0)];
    const ChanParams &p = a.params[ch];
    ChanState *st = a.state + ch;
    const int gq = active ? g : 0;          // queue column read by idle lanes (any valid one)

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
        for (int t = 0; t < niter; t++) {
            const int c = t - IT_LAT;
            if (c >= 0 && c < nsteps) {
                const float *in = sm.aud[c & 1];
                float *out = sm.lat + (c % 5) * ND * SMS;
#pragma unroll 1
                for (int i0 = 0; i0 < ND; i0 += 8) {
                    // the 8 inputs first: their shared-memory latency is paid once, not once per sample
                    float xin[8], yo[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) xin[i] = in[(i0 + i) * SMS + gq];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        float f = xin[i], acc = 0.0f, fn = f;
#pragma unroll
                        for (int j = 0; j < 10; j++) {
                            const float gg = s[j];
                            fn = fmaf(-k[j], gg, f);
                            const float gn = fmaf(fn, k[j], gg);
                            acc = fmaf(gn, v[j], acc);
                            if (j > 0) s[j - 1] = gn;
                            f = fn;
                        }
                        yo[i] = fmaf(fn, v[10], acc);
                        s[9] = fn;
                    }
                    if (active) {
#pragma unroll
                        for (int i = 0; i < 8; i++) out[(i0 + i) * SMS + g] = yo[i];
                    }
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
        // ---- WDSP AGC (audio_agc.c:349-595), mono, 12 ksps, 49-sample look-ahead: the detector.  It turns the lattice
        // output x[n] (and the delayed sample x[n-49] for the back-averages) into the control voltage "volts" per
        // sample; the gain law and the multiplication of x[n-49] run one pipeline stage later (BQ warp).
        const AgcP ap = p.agc;
        AgcRun ar = { 0, 0, st->agc_ring_max, st->agc_volts, st->agc_save_volts, st->agc_fast_backaverage,
                      st->agc_hang_backaverage, st->agc_hang_counter, st->agc_decay_type, st->agc_state,
                      st->agc_action, st->agc_hang_action };
        float *latp = sm.lat + g;
        // history x[-49..-1] from the 192-slot state ring -> rows of steps -1 and -2 of the lattice-output ring
        const int in_index = st->agc_in_index;
        if (active) {
            for (int kk = 1; kk <= AGC_W; kk++) {
                int idx = in_index - (kk - 1);
                idx %= AGC_RB; if (idx < 0) idx += AGC_RB;
                latp[(LR - kk) * SMS] = st->agc_ring[idx];
            }
        }
        // Sliding maximum of |x| over the newest 49 samples (audio_agc.c:409-429 rescans on demand): van Herk /
        // Gil-Werman decomposition with blocks = steps of 32, see rx_ssb_fused.cu.
        int s1 = 0, s2 = 1;
        if (active) {
            float m = 0.0f;
            for (int o = ND - 1; o >= 0; o--) {
                m = fmaxf(m, fabsf(latp[(LR - ND + o) * SMS]));
                sm.smax[s1][o * SMS + g] = m;
            }
            m = 0.0f;
            for (int o = ND - 1; o >= 16; o--) {
                m = fmaxf(m, fabsf(latp[(LR - 2 * ND + o) * SMS]));
                sm.smax[s2][o * SMS + g] = m;
            }
        }
        __syncwarp();
        const bool any_hang = __any_sync(0xffffffffu, active && (ap.hang_enable || ar.state == 2 || ar.state == 4 || ar.decay_type != 0 || ar.hang_counter > 0));
        for (int t = 0; t < niter; t++) {
            const int c = t - IT_AGC;
            if (c >= 0 && c < nsteps && active && ap.mode != 5) {
                const int row0 = (c % 5) * ND;                  // ring row of the first sample of this step
                const float *in = latp + row0 * SMS;
                float *out = sm.agc[c & 1] + g;
                const float *S1 = sm.smax[s1] + g, *S2 = sm.smax[s2] + g;
                const float mprev = S1[0];
                float pmax = 0.0f;
                auto detect = [&](auto hangc) {
                    constexpr bool HANG = decltype(hangc)::value;
#pragma unroll 1
                for (int k8 = 0; k8 < ND; k8 += AG) {
                    // all operands of the AG samples first.  The delayed sample x[n-49] of group element j sits at ring
                    // row (row0 + k8 + j - 49) mod LR; a group wraps at most between its first and second element.
                    int ra = row0 + k8 - AGC_W; if (ra < 0) ra += LR;
                    int rb = ra + 1; if (rb >= LR) rb -= LR;
                    const float *dA = latp + ra * SMS, *dB = latp + rb * SMS;
                    const bool two = k8 < 16;                   // window still reaches into the chunk before the previous one
                    const float *pc = two ? S2 + (16 + k8) * SMS : S1 + (k8 - 16) * SMS;
                    const float *pin = in + k8 * SMS;
                    float x[AG], dly[AG], cmx[AG];
#pragma unroll
                    for (int j = 0; j < AG; j++) {
                        x[j] = pin[j * SMS];
                        dly[j] = (j == 0) ? dA[0] : dB[(j - 1) * SMS];
                        const float sfx = pc[j * SMS];
                        cmx[j] = two ? fmaxf(mprev, sfx) : sfx;
                    }
#pragma unroll
                    for (int j = 0; j < AG; j++) {
                        const float abs_out = fabsf(dly[j]), abs_in = fabsf(x[j]);
                        pmax = fmaxf(pmax, abs_in);
                        ar.fast_backaverage = fmaf(ap.fast_backmult, abs_out, __fmul_rn(ap.onemfast_backmult, ar.fast_backaverage));
                        ar.hang_backaverage = fmaf(ap.hang_backmult, abs_out, __fmul_rn(ap.onemhang_backmult, ar.hang_backaverage));
                        ar.ring_max = fmaxf(pmax, cmx[j]);
                        const float dv = __fsub_rn(ar.ring_max, ar.volts);
                        const bool attack = ar.ring_max >= ar.volts;
                        float mult_sel = ap.attack_mult;
                        bool upd = true;
                        int nstate = ar.state;
                        if constexpr (!HANG) {
                            // hang AGC disabled on every channel of this warp (the default, ui_configuration.c:81): only
                            // states 0 / 1 / 3 occur and the 5-state machine reduces to selects
                            const bool fast = (ar.state == 0) ? (ar.volts > __fmul_rn(ap.pop_ratio, ar.fast_backaverage))
                                                              : ((ar.state == 1) && (ar.volts > ar.save_volts));
                            if (attack && ar.state >= 2) ar.save_volts = ar.volts;
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
                        if (upd) ar.volts = fmaf(dv, mult_sel, ar.volts);
                        if (ar.volts < ap.min_volts) { ar.volts = ap.min_volts; ar.action = 0; } else { ar.action = 1; }
                        out[(k8 + j) * SMS] = ar.volts;
                    }
                }
                };
                if (any_hang) detect(std::true_type{}); else detect(std::false_type{});
                ar.hang_action = (ar.hang_backaverage > ap.hang_level) ? 1 : 0;
                {
                    // suffix maxima of this step replace those of the step before the previous one; roles rotate
                    float *Sn = sm.smax[s2] + g;
                    float m = 0.0f;
#pragma unroll 1
                    for (int o8 = ND - 8; o8 >= 0; o8 -= 8) {
                        float rv[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) rv[j] = in[(o8 + j) * SMS];
#pragma unroll
                        for (int j = 7; j >= 0; j--) { m = fmaxf(m, fabsf(rv[j])); Sn[(o8 + j) * SMS] = m; }
                    }
                    const int tmp = s1; s1 = s2; s2 = tmp;
                }
            }
            __syncthreads();
        }
        if (active && ap.mode != 5) {
            const long long T = (long long)nsteps * ND;
            int new_in = (int)(((long long)st->agc_in_index + T) % AGC_RB);
            int new_out = (int)((((long long)st->agc_out_index + T) % AGC_RB + AGC_RB) % AGC_RB);
            const int rend = (int)(T % LR);                      // ring row one past the newest sample
            for (int kk = 1; kk <= AGC_W; kk++) {
                int idx = new_in - (kk - 1);
                idx %= AGC_RB; if (idx < 0) idx += AGC_RB;
                int rr = rend - kk; if (rr < 0) rr += LR;
                st->agc_ring[idx] = latp[rr * SMS];
            }
            st->agc_in_index = new_in; st->agc_out_index = new_out;
            st->agc_ring_max = ar.ring_max; st->agc_volts = ar.volts; st->agc_save_volts = ar.save_volts;
            st->agc_fast_backaverage = ar.fast_backaverage; st->agc_hang_backaverage = ar.hang_backaverage;
            st->agc_hang_counter = ar.hang_counter; st->agc_decay_type = ar.decay_type; st->agc_state = ar.state;
            st->agc_action = ar.action; st->agc_hang_action = ar.hang_action;
        }
        return;
    }

    if (warp == W_BQ) {
        // ---- AGC gain law on the delayed sample (audio_agc.c:563-570; mode 5 = fixed gain :354-365), then the
        // fixed gain (:2513-2524) and the 4-stage DF1 cascade IIR_biquad_1 (:2527) at 12 ksps.  Per stage the terms
        // that do not depend on the newest input are summed ahead of time (t), so the sample-to-sample critical path is
        // one FMA per stage.  Stages whose coefficients are {1,0,0,0,0} on every channel of the CTA (notch / peak off:
        // the default) are skipped; their state is the last two samples that went through.
        float bc[4][5]; BiquadS bs[4]; float tq[4];
        unsigned skipmask = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
#pragma unroll
            for (int q = 0; q < 5; q++) bc[s][q] = p.bq1[s][q];
            bs[s] = st->bq1[s];
            const bool ident = bc[s][0] == 1.0f && bc[s][1] == 0.0f && bc[s][2] == 0.0f && bc[s][3] == 0.0f && bc[s][4] == 0.0f;
            if (__all_sync(0xffffffffu, ident || !active)) skipmask |= 1u << s;
            tq[s] = fmaf(bc[s][3], bs[s].y1, fmaf(bc[s][1], bs[s].x1, fmaf(bc[s][2], bs[s].x2, __fmul_rn(bc[s][4], bs[s].y2))));   // same order as in the loop
        }
        const float scale_gain = p.scale_gain;
        const AgcP ap = p.agc;
        const bool agc_off = ap.mode == 5;
        const float *latp = sm.lat + g;
        float xl1 = 0.0f, xl2 = 0.0f;        // the last two cascade inputs (state of skipped leading stages)
        // one step of 32 samples with the skipped stages known at compile time
        auto run_step = [&](auto maskc, const float *in, float *out, int row0) {
            constexpr unsigned MASK = decltype(maskc)::value;
#pragma unroll 1
            for (int i0 = 0; i0 < ND; i0 += 8) {
                // volts of the 8 samples and the delayed samples x[n-49] (AGC off: the current samples x[n])
                int ra = row0 + i0 - (agc_off ? 0 : AGC_W); if (ra < 0) ra += LR;
                int rb = ra + 1; if (rb >= LR) rb -= LR;
                const float *dA = latp + ra * SMS, *dB = latp + rb * SMS;
                float xv[8], vv[8];
#pragma unroll
                for (int i = 0; i < 8; i++) { vv[i] = in[(i0 + i) * SMS]; xv[i] = (i == 0) ? dA[0] : dB[(i - 1) * SMS]; }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    // Math_log10f_fast (uhsdr_math.c:27-41) of inv_max_input * volts (> 0): exponent / mantissa by bit
                    // operations, the cubic in Horner form
                    const unsigned ub = __float_as_uint(__fmul_rn(ap.inv_max_input, vv[i]));
                    const float F = __uint_as_float((ub & 0x007fffffu) | 0x3f000000u);      // frexpf mantissa in [0.5, 1)
                    const float E = (float)((int)(ub >> 23) - 126);
                    float Y = fmaf(1.23149591368684f, F, -4.11852516267426f);
                    Y = fmaf(Y, F, 6.02197014179219f);
                    Y = fmaf(Y, F, -3.13396450166353f);
                    float vo = __fmul_rn(__fadd_rn(Y, E), 0.3010299956639812f);
                    vo = fminf(vo, 0.0f);
                    const float mult = agc_off ? ap.fixed_gain : __fdividef(fmaf(-ap.slope_constant, vo, ap.out_target), vv[i]);
                    xv[i] = __fmul_rn(xv[i], mult);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float x = __fmul_rn(xv[i], scale_gain);
                    if (i == 6) xl2 = x;
                    if (i == 7) xl1 = x;
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        if (!(MASK & (1u << s))) {
                            const float w = fmaf(bc[s][2], bs[s].x1, __fmul_rn(bc[s][4], bs[s].y1));     // next sample's x2 / y2 terms
                            const float y = fmaf(bc[s][0], x, tq[s]);
                            tq[s] = fmaf(bc[s][3], y, fmaf(bc[s][1], x, w));
                            bs[s].x2 = bs[s].x1; bs[s].x1 = x; bs[s].y2 = bs[s].y1; bs[s].y1 = y;
                            x = y;
                        }
                    }
                    xv[i] = x;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) out[(i0 + i) * SMS] = xv[i];
            }
        };
        if (skipmask != 0xbu) skipmask = 0;      // only the default plan (bass shelf alone) has a specialised loop
        for (int t = 0; t < niter; t++) {
            const int c = t - IT_BQ;
            if (c >= 0 && c < nsteps && active) {
                const float *in = sm.agc[c & 1] + g;
                float *out = sm.bq[c & 1] + g;
                if (skipmask == 0xbu) run_step(std::integral_constant<unsigned, 0xbu>{}, in, out, (c % 5) * ND);
                else run_step(std::integral_constant<unsigned, 0u>{}, in, out, (c % 5) * ND);
            }
            __syncthreads();
        }
        if (active) {
            // a skipped (pass-through) stage saw the output of the nearest computed stage before it
            float s1v = xl1, s2v = xl2;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if (skipmask & (1u << s)) { bs[s].x1 = s1v; bs[s].x2 = s2v; bs[s].y1 = s1v; bs[s].y2 = s2v; }
                else { s1v = bs[s].y1; s2v = bs[s].y2; }
                st->bq1[s] = bs[s];
            }
        }
        return;
    }

    // warp == W_POST: x4 interpolator (:2560-2577), anti-alias lattice (:2581-2583, 6 stages, some paths), treble biquad
    // (:2832), x10, output formatting (:2845-2941), straight to global memory: 32 bytes (4 output samples) per channel
    // and decimated sample.
    {
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
        float ak[6], av[7], as[6];
        const int n = p.aa.n;
#pragma unroll
        for (int j = 0; j < 6; j++) {
            ak[j] = (n == 6) ? __ldg(pool + p.aa.k_off + j) : 0.0f;
            av[j] = (n == 6) ? __ldg(pool + p.aa.v_off + j) : 0.0f;
            as[j] = (n == 6) ? st->aa_s[j] : 0.0f;
        }
        av[6] = (n == 6) ? __ldg(pool + p.aa.v_off + 6) : 1.0f;
        const bool any_aa = __any_sync(0xffffffffu, active && n == 6);
        float tc[5];
#pragma unroll
        for (int q = 0; q < 5; q++) tc[q] = p.bq2[q];
        BiquadS ts = st->bq2;
        float tt = fmaf(tc[3], ts.y1, fmaf(tc[1], ts.x1, fmaf(tc[2], ts.x2, __fmul_rn(tc[4], ts.y2))));   // same order as in the loop
        const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
        int4 *__restrict__ dst = reinterpret_cast<int4 *>(reinterpret_cast<int2 *>(a.audio) + chan_base);
        float4 *__restrict__ dst_f = a.audio_f ? reinterpret_cast<float4 *>(a.audio_f + chan_base) : nullptr;
        const uint8_t *__restrict__ mute = a.mute ? a.mute + (size_t)ch * (size_t)a.mute_stride : nullptr;

        // A 0 dB shelf has b0 = 1, b1 = -a1, b2 = -a2 (audio_driver.c:906-964 with A = 1): with consistent state
        // (y1 = x1, y2 = x2) the stage is the identity.  Then it is not computed (the reference's own result differs
        // from its input by float rounding only); the state follows the signal so that later launches agree.
        const bool tr_unity = __all_sync(0xffffffffu, !active || (tc[0] == 1.0f && tc[1] == -tc[3] && tc[2] == -tc[4] &&
                                                                  ts.x1 == ts.y1 && ts.x2 == ts.y2));
        // one 32-sample block (8 decimated samples -> 32 outputs = 256 bytes of the channel's row)
        auto run_block = [&](auto aac, auto trc, const float *in, int4 *d4, float4 *df, bool muted) {
            constexpr bool AA = decltype(aac)::value, TR = decltype(trc)::value;
            const int mm = muted ? 0 : -1;                // external_mute: zeros out, all state advanced (:2845-2853)
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                float xv[4];
#pragma unroll
                for (int i = 0; i < 4; i++) xv[i] = in[(4 * h + i) * SMS];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float x = xv[i];
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float pre = fmaf(ih[2], ic[j][2], fmaf(ih[1], ic[j][1], __fmul_rn(ih[0], ic[j][0])));
                        o[j] = fmaf(x, ic[j][3], pre);
                    }
                    ih[0] = ih[1]; ih[1] = ih[2]; ih[2] = x;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float y = o[j];
                        if constexpr (AA) {
                            float f = y, acc = 0.0f, fn = y;
#pragma unroll
                            for (int q = 0; q < 6; q++) {
                                const float gg = as[q];
                                fn = fmaf(-ak[q], gg, f);
                                const float gn = fmaf(fn, ak[q], gg);
                                acc = fmaf(gn, av[q], acc);
                                if (q > 0) as[q - 1] = gn;
                                f = fn;
                            }
                            acc = fmaf(fn, av[6], acc);
                            as[5] = fn;
                            y = (n == 6) ? acc : y;
                        }
                        float z = y;
                        if constexpr (TR) {
                            const float w = fmaf(tc[2], ts.x1, __fmul_rn(tc[4], ts.y1));
                            z = fmaf(tc[0], y, tt);
                            tt = fmaf(tc[3], z, fmaf(tc[1], y, w));
                            ts.x2 = ts.x1; ts.x1 = y; ts.y2 = ts.y1; ts.y1 = z;
                        }
                        o[j] = z;
                    }
                    if constexpr (!TR) { ts.x1 = o[3]; ts.y1 = o[3]; ts.x2 = o[2]; ts.y2 = o[2]; }
                    const int pos = 4 * h + i;
                    const int w0 = format_audio_word(__fmul_rn(o[0], 10.0f)) & mm, w1 = format_audio_word(__fmul_rn(o[1], 10.0f)) & mm;   // LINE_OUT_SCALING_FACTOR (:2860)
                    const int w2 = format_audio_word(__fmul_rn(o[2], 10.0f)) & mm, w3 = format_audio_word(__fmul_rn(o[3], 10.0f)) & mm;
                    // the four output samples {l, r} x 4 = 32 bytes: one 256-bit store (STG.E.ENL2.256)
                    asm volatile("st.global.v8.b32 [%0], {%1, %1, %2, %2, %3, %3, %4, %4};" ::"l"(d4 + 2 * pos), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
                    if (df) df[pos] = muted ? make_float4(0.0f, 0.0f, 0.0f, 0.0f)
                                            : make_float4(__fmul_rn(o[0], 10.0f), __fmul_rn(o[1], 10.0f), __fmul_rn(o[2], 10.0f), __fmul_rn(o[3], 10.0f));
                }
            }
        };
        for (int t = 0; t < niter; t++) {
            const int c = t - IT_POST;
            if (c >= 0 && c < nsteps && active) {
                const float *in = sm.bq[c & 1] + g;
#pragma unroll 1
                for (int blk = 0; blk < 4; blk++) {
                    const bool muted = mute && mute[c * 4 + blk];
                    int4 *d4 = dst + (size_t)c * 64 + blk * 16;
                    float4 *df = dst_f ? dst_f + (size_t)c * 32 + blk * 8 : nullptr;
                    if (any_aa) run_block(std::true_type{}, std::true_type{}, in + blk * 8 * SMS, d4, df, muted);
                    else if (tr_unity) run_block(std::false_type{}, std::false_type{}, in + blk * 8 * SMS, d4, df, muted);
                    else run_block(std::false_type{}, std::true_type{}, in + blk * 8 * SMS, d4, df, muted);
                }
            }
            __syncthreads();
        }
        if (active) {
            for (int q = 0; q < INTERP_HIST - 3; q++) st->interp_hist[q] = 0.0f;
#pragma unroll
            for (int q = 0; q < 3; q++) st->interp_hist[INTERP_HIST - 3 + q] = ih[q];
            if (n == 6) {
#pragma unroll
                for (int j = 0; j < 6; j++) st->aa_s[j] = as[j];
            }
            st->bq2 = ts;
        }
    }
}

cudaError_t launch_rx_ssb_tc(const RxArgs &a, const FusedCoefs &fc, int hil_ci, int hil_cq, int sm_count, cudaStream_t stream)
{
    if (a.num_items <= 0) return cudaSuccess;
    if (a.nblocks % 4 != 0 || a.chan_list == nullptr) return cudaErrorInvalidValue;
    int per = (a.num_items + sm_count - 1) / sm_count;
    per = ((per + 3) / 4) * 4;
    if (per > FG) per = FG;
    if (per < 4) per = 4;
    const int grid = (a.num_items + per - 1) / per;
    cudaError_t e = cudaFuncSetAttribute(rx_ssb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    rx_ssb_tc_kernel<<<grid, NTHREADS, sizeof(Smem), stream>>>(a, fc, per, hil_ci, hil_cq);
    return cudaGetLastError();
}

bool rx_ssb_tc_available() { return true; }

}  // namespace uhsdr

#else   // UHSDR_EXACT: reference-order arithmetic has no tensor-core form

namespace uhsdr {
cudaError_t launch_rx_ssb_tc(const RxArgs &, const FusedCoefs &, int, int, int, cudaStream_t) { return cudaErrorNotSupported; }
bool rx_ssb_tc_available() { return false; }
}  // namespace uhsdr

#endif
