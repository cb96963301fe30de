// rx_ssb_tc.cu -- fused narrow-SSB/CW receiver kernel with BOTH FIR stages on the 5th-generation tensor
// cores (tcgen05.mma, accumulators in TMEM).  Shipping build only; the exact build keeps the CUDA-core
// kernel of rx_ssb_fused.cu.
//
// Chain (FilterPathInfo[4..47], mchf-eclipse/drivers/audio/audio_filter.c:147-922; flow in
// audio_driver.c:2603-2942):
//   format + IQ correction + Fs/4 translate                                         (front-end warps, FP32 pipe)
//   -> 83-tap /4 decimator on I and Q                                               (tensor cores)
//   -> 199-tap Hilbert pair @12 ksps -> I +/- Q                                     (tensor cores)
//   -> 10-stage lattice IIR | WDSP AGC | gain, 4-stage biquad | x4 interpolator, (anti-alias lattice),
//      treble biquad, x10, int32 << 16                                              (four serial warps)
//
// One persistent CTA owns up to 28 channels for the whole launch; every sample crosses HBM once.
//
// Both FIRs are Toeplitz GEMMs (north_star form (1)) with the taps as the A operand and the samples of all
// channels of the CTA as the B operand, accumulated INCREMENTALLY: every 128-sample step brings one slab of
// new samples into a small shared-memory ring, and that slab is multiplied into every output chunk it
// contributes to (the TMEM accumulators of the chunks stay live across steps), so shared memory holds two
// steps of samples instead of a whole filter window.
//   decimator (arm_fir_decimate_f32.c:455-486): y[m] = sum_k c[k] x[4m - 82 + k].  Output chunk = 128 outputs
//     = 512 inputs; with window slot p = input index - (512 chunk - 96):  D[m][col] = sum_p T[m][p] X[p][col],
//     T[m][p] = c[p - 4m - 14], 38 k-steps of 16.  The 128 x 16 slab of k-step kk is rows [148 - 4kk, +128) of
//     ONE table G[r][q] = c[q - 4r + 578] stored with linear rows, so the A descriptor slides 64 B per k-step.
//     I and Q share the taps: columns = 32 I channels | 32 Q channels (N = 64).
//   Hilbert pair (arm_fir_f32.c:522-529): y[n] = sum_k c[k] d[n - 198 + k].  Output chunk = 128 outputs, window
//     slot p = index - (128 chunk - 208), T[m][p] = c[p - m - 10], 21 k-steps; slab kk = rows [320 - 16kk, +128)
//     of G[r][q] = c[q - r + 310].  I and Q accumulate into the same D (USB: I + Q; LSB: Q stored negated).
// Arithmetic: BF16 operands, FP32 accumulation, samples rounded to 16 significant bits and split exactly
// x = x1 + x2, taps c = c1 + c2, D = c1 x1 + c2 x1 + c1 x2 -- relative error ~5e-6 per stage (106 dB), measured
// in scripts/micro/umma_toeplitz.cu and umma_dec_toeplitz.cu.  The rounded values are also what the channel
// state keeps, so splitting a call in two reproduces the one-call result bit for bit.
// tcgen05.mma time here is set by operand fetch from shared memory (~128 B/clk: 48 cycles for the 128x64x16
// decimator MMA, 40 for the 128x32x16 Hilbert MMA); the MMAs are issued under elect.sync so that ptxas emits
// them back to back (issued from an `if (lane == 0)` region each one costs 56 cycles of a serialisation loop).
#include <cuda_bf16.h>

#include <type_traits>

#include "dsp_device.cuh"
#include "kernels.h"
#include "uhsdr_b200.h"

#if !UHSDR_EXACT

namespace uhsdr {

namespace {

constexpr int FG = 28;             // channel slots per CTA (7 front-end warps x 4 channels)
constexpr int CH4 = 128;           // input samples per step (4 blocks)
constexpr int ND = 32;             // decimated samples per step
constexpr int SMS = 29;            // channel-minor stride of the serial-stage queues (odd: conflict free)
constexpr int AGC_W = 49;          // attack_buffsize at 12 ksps (audio_agc.c:290)
constexpr int LR = 5 * ND;          // rows of the lattice-output ring (decimated samples)
constexpr int AG = 2;              // AGC samples per loop iteration (short body: level-0 instruction cache)
constexpr int NWARP_FE = FG / 4;
// tensor-core FIRs
constexpr int V = 8;               // virtual steps in front of the first real one: they carry the filter histories in
constexpr int XSLOTS = 2 * CH4;    // decimator input ring: two steps
constexpr int XLBO = 144;          // byte stride between time groups: 128-byte core matrix + 16 bytes, so that the front-end's 8-byte stores of a
                                   // half-warp (4 channels x two halves of a row, in two time groups 4 apart) tile one 128-byte bank row
constexpr int XSBO = (XSLOTS / 8) * XLBO;
constexpr int XR_BYTES = 8 * XSBO;                 // one array: [I groups 0..3 | Q groups 0..3][time/8][channel%8][time%8] bf16
constexpr int HSLOTS = 2 * ND;     // Hilbert input ring: two steps
constexpr int HLBO = 144;          // time-group stride: core matrix + 16 bytes (the epilogue stores of 32 consecutive slots hit 4 distinct bank groups)
constexpr int HSBO = (HSLOTS / 8) * HLBO;
constexpr int HR_BYTES = 4 * HSBO;                 // one array: [channel group 0..3][time/8][channel%8][time%8] bf16
constexpr int DROWS = 276, DR0 = 148, DC0 = 578, DPLANE = DROWS * 16;    // decimator Toeplitz table (see above)
constexpr int HROWS = 448, HR0 = 320, HC0 = 310, HPLANE = HROWS * 16;    // Hilbert Toeplitz table
constexpr int MMA_PAUSE = 60;         // cycles between k-steps of the MMA issue (sweeps: profiles/r01_tc2_experiments.txt, r02_tc_experiments.txt)
// 2 decimator accumulators of 128 columns (c1 x1 + c2 x1 for I | Q in 0..63, c1 x2 in 64..127), 4 Hilbert accumulators of 64
// (c1 x1 + c2 x1 in 0..31, c1 x2 in 32..63): the c1 products of both halves of the split come from ONE MMA with twice the columns,
// so the Toeplitz slab -- 4 KB of the 5-6 KB an MMA fetches from shared memory -- is read twice per k-step instead of three times
constexpr int DEC_COL0 = 0, DEC_COLS = 128, HIL_COL0 = 256, HIL_COLS = 64, TMEM_COLS = 512;
// warp roles (warp id % 4 is the scheduler and the TMEM lane quadrant)
// scheduler 0: front end 0, 4 | epilogue 0 | biquads (+ AGC helper) | MMA issue     1: front end 1, 5 | epilogue 1 | output B
//           2: front end 2, 6 | epilogue 2 | lattice                                 3: front end 3 | output A | epilogue 3 | AGC detector
// (17 warps: scheduler 0 has five; it gets the two lightest roles, and the AGC detector -- the longest recurrence -- the emptiest one)
constexpr int W_POSTA = NWARP_FE, W_EPI = NWARP_FE + 1, W_BQ = NWARP_FE + 5, W_POSTB = NWARP_FE + 6, W_LAT = NWARP_FE + 7, W_AGC = NWARP_FE + 8,
              W_MMA = NWARP_FE + 9;
constexpr int NTHREADS = 32 * (NWARP_FE + 10);
// software pipeline, in steps of 128 input samples.  Step s (virtual steps included) is written into the decimator
// ring at iteration s, its decimator MMAs are issued at s + 1, the decimator outputs leave TMEM for the Hilbert ring
// at s + 2, the Hilbert MMAs are issued at s + 3, the Hilbert outputs leave TMEM at s + 4, then lattice s + 5,
// AGC detector s + 6, AGC gain law s + 7 (on the two epilogue warps that are idle in that iteration), biquad cascade s + 8,
// interpolator / treble / output formatting s + 9.
constexpr int IT_LAT = V + 5, IT_AGC = V + 6, IT_GAIN = V + 7, IT_BQ = V + 8, IT_POST = V + 9;
constexpr int PIPE_DEPTH = IT_POST;

// Per-role timing (tools only, -DUHSDR_TC_PROF: `make prof`, scripts/tc_role_times.py): every warp of CTA 0 adds up the
// cycles between the top of a pipeline iteration and its arrival at the step barrier, for the iterations in which all
// roles have work (steady state); [warp][0] = work cycles, [1] = iterations counted, [2] = whole iterations (barrier to barrier).
#ifdef UHSDR_TC_PROF
__device__ unsigned long long g_tc_prof[32][4];
__device__ int g_tc_pause = -1;  // MMA issue pause override (cycles), -1 = MMA_PAUSE
__device__ int g_tc_knock;      // knock-out mask (timing experiments only, results wrong): 1 front end, 2 MMAs, 4 decimator epilogue,
                                // 8 Hilbert epilogue, 16 gain law, 32 lattice, 64 AGC detector, 128 biquads, 256 output
#define KNOCK(b) ((knock & (b)) != 0)
#define PROF_DECL long long prof_t0 = 0, prof_work = 0, prof_tot = 0, prof_last = 0; int prof_n = 0
#define PROF_TOP(it) do { prof_t0 = clock64(); if ((it) > PIPE_DEPTH + 1 && (it) < nsteps - 1 && prof_last) prof_tot += prof_t0 - prof_last; prof_last = prof_t0; } while (0)
#define PROF_END(it) do { if ((it) > PIPE_DEPTH && (it) < nsteps - 1) { prof_work += clock64() - prof_t0; prof_n++; } } while (0)
#define PROF_SAVE() do { if (blockIdx.x == 0 && lane == 0) { g_tc_prof[warp][0] = (unsigned long long)prof_work; g_tc_prof[warp][1] = (unsigned long long)prof_n; g_tc_prof[warp][2] = (unsigned long long)prof_tot; } } while (0)
#else
#define KNOCK(b) false
#define PROF_DECL
#define PROF_TOP(it)
#define PROF_END(it)
#define PROF_SAVE()
#endif

struct Smem {
    alignas(128) unsigned char xring[2][XR_BYTES];    // [I1 | Q1], [I2 | Q2]: bf16 split of the front-end outputs
    alignas(128) unsigned char hring[4][HR_BYTES];    // I1, I2, Q1, Q2: bf16 split of the decimator outputs
    alignas(128) unsigned char gh[4][2 * HPLANE];     // Hilbert Toeplitz tables: hil_i c1, c2, hil_q c1, c2
    alignas(128) unsigned char gd[2][2 * DPLANE];     // decimator Toeplitz tables: c1, c2
    float aud[2][ND * SMS];
    float lat[LR * SMS];              // lattice output, a ring of 5 steps: AGC detector and gain stage read x[n-49] from it
    float agc[2][ND * SMS];           // AGC "volts" per sample (detector -> gain stage)
    float gq[2][ND * SMS];            // delayed sample x gain (gain stage -> biquad cascade)
    float bq[2][(3 + ND) * SMS];      // biquad output: rows 0..2 = the last three samples of the previous step (interpolator history), then the step
    float smax[3][ND * SMS];          // suffix maxima of |x| per step (van Herk / Gil-Werman), step c in buffer c % 3
    int chan[32];                     // channel index of every slot (epilogue warps: state save / restore)
    alignas(8) unsigned long long bar_dec[2], bar_hil[2];
    unsigned tmem_base;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

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

__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B; lbo = byte distance
// between the two 16-byte K halves of a k-step, sbo = byte distance between 8-row groups.  The start address
// sits in the low 14 bits (units of 16 B), so moving an operand is an addition to the descriptor.
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr, unsigned lbo, unsigned sbo)
{
    return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// the three products of one k-step in two MMAs: D[0 .. 2N) (+)= A1 [B1 | B2] (the lo half B2 of the sample split lies right behind
// the hi half B1 in shared memory: one descriptor, idesc2 = twice the columns); D[0 .. N) += A2 B1.  The epilogue adds the halves.
__device__ __forceinline__ void umma2_bf16(unsigned tmem_d, unsigned long long a1, unsigned long long a2, unsigned long long b1,
                                           unsigned idesc2, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 q, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %4, p;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, q;\n\t}"
                 ::"r"(tmem_d), "l"(a1), "l"(a2), "l"(b1), "r"(idesc2), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// x -> two bf16 values: hi = round-to-nearest bf16 of x, lo = the exact remainder x - hi truncated to bf16 (8 more
// significant bits, so hi + lo carries 16-17 bits of x).  The split is idempotent: splitting h = hi + lo again gives
// (hi, lo), and h is what the channel state keeps, so cutting a call in two reproduces the one-call result bit for bit.
__device__ __forceinline__ void split_bf16(float x, unsigned &hi, unsigned &lo)
{
    unsigned short hb;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(hb) : "f"(x));
    hi = hb;
    lo = __float_as_uint(x - __uint_as_float(hi << 16)) >> 16;
}
// two samples -> one word of the hi array and one of the lo array (element 0 in the low half): one F2FP for both hi
__device__ __forceinline__ void split_pack2(float x0, float x1, unsigned &whi, unsigned &wlo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(whi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(whi << 16), r1 = x1 - __uint_as_float(whi & 0xffff0000u);
    wlo = __byte_perm(__float_as_uint(r0), __float_as_uint(r1), 0x7632);
}
__device__ __forceinline__ float join_bf16(unsigned hi, unsigned lo) { return __uint_as_float(hi << 16) + __uint_as_float(lo << 16); }

struct FirLaneState {
    float te1, te2, te3;     // teta*_old
    float c1, c2;            // M_c1, M_c2
    int clip;                // bit0 quarter, bit1 half, bit2 full
};

#define TMEM_LD_X32(v, taddr)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                 \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27," \
                 "%28,%29,%30,%31}, [%32];"                                                                                \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),         \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),   \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])  \
                 : "r"(taddr))

#define TMEM_LD_X16(v, taddr)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"        \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                                        \
                 : "r"(taddr))
#define TMEM_LD_X4(v, taddr) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr))

}  // namespace

__global__ void __launch_bounds__(NTHREADS, 1)
rx_ssb_tc_kernel(const __grid_constant__ RxArgs a, int chans_per_cta, int dec_c, int hil_ci, int hil_cq)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cta_first = blockIdx.x * chans_per_cta;
    const int n_here = min(chans_per_cta, a.num_items - cta_first);
    const int nsteps = a.nblocks / 4;
    if (nsteps <= 0) return;
    const int niter = nsteps + PIPE_DEPTH;
    const int s_end = V + nsteps;                      // one past the last real step
    const float *__restrict__ pool = a.pool;
#ifdef UHSDR_TC_PROF
    const int knock = g_tc_knock;
#endif

    // ---- one-time setup by all threads: Toeplitz tables, zeroed rings, barriers, TMEM ----
    for (int i = threadIdx.x; i < 2 * HROWS * 16; i += NTHREADS) {
        const int which = i / (HROWS * 16), e = i % (HROWS * 16);
        const int r = e >> 4, q = e & 15;
        const int kidx = q - r + HC0;
        const float cv = (kidx >= 0 && kidx < 199) ? __ldg(pool + (which ? hil_cq : hil_ci) + kidx) : 0.0f;
        const __nv_bfloat16 c1 = __float2bfloat16_rn(cv);
        const __nv_bfloat16 c2 = __float2bfloat16_rn(cv - __bfloat162float(c1));
        const int off = (q >> 3) * HPLANE + r * 16 + (q & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(sm.gh[2 * which] + off) = c1;
        *reinterpret_cast<__nv_bfloat16 *>(sm.gh[2 * which + 1] + off) = c2;
    }
    for (int i = threadIdx.x; i < DROWS * 16; i += NTHREADS) {
        const int r = i >> 4, q = i & 15;
        const int kidx = q - 4 * r + DC0;
        const float cv = (kidx >= 0 && kidx < 83) ? __ldg(pool + dec_c + kidx) : 0.0f;
        const __nv_bfloat16 c1 = __float2bfloat16_rn(cv);
        const __nv_bfloat16 c2 = __float2bfloat16_rn(cv - __bfloat162float(c1));
        const int off = (q >> 3) * DPLANE + r * 16 + (q & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(sm.gd[0] + off) = c1;
        *reinterpret_cast<__nv_bfloat16 *>(sm.gd[1] + off) = c2;
    }
    for (int i = threadIdx.x; i < 2 * XR_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm.xring)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 4 * HR_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm.hring)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) sm.chan[threadIdx.x] = a.chan_list[cta_first + (threadIdx.x < n_here ? threadIdx.x : 0)];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_dec[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_dec[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_hil[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_hil[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // tables and zeroed rings -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;

    if (warp < NWARP_FE) {
        // ======================= front-end warp: 4 channels x 8 lanes ============================
        // A 128-sample step is done in two passes of 64 samples (two 32-sample blocks); in a pass lane slot r owns the 8
        // consecutive samples 8r .. 8r+7, i.e. a quarter of block r >> 2.  (Two short passes instead of one long one: half the
        // instructions in the loop body -- the instruction cache is the scarce resource of this kernel.)
        // A pass is two 32-sample blocks; lane = cl + 4 rp with rp = q_lo + 2 h + 4 q_hi: the lane works on block h and owns its
        // samples 16 j + 4 q .. + 3 (q = q_lo + 2 q_hi; j = 0, 1: one 32-byte load each).  The 8 lanes of a channel then read 256
        // contiguous bytes per load instruction (two full 128-byte lines instead of eight quarter-used ones: the load / store
        // data path is the busiest unit of this kernel), and the 16 lanes of a half-warp (q_hi fixed) store 4 channels x two
        // row halves in two time groups 4 apart = one conflict-free 128-byte bank row of the ring (time-group stride 144 B).
        const int cl = lane & 3, rp = lane >> 2;
        const int q_lo = rp & 1, hb = (rp >> 1) & 1, q_hi = rp >> 2;
        const int g = warp * 4 + cl;                   // channel slot in the CTA
        const bool active = g < n_here;
        const int ch = active ? a.chan_list[cta_first + g] : a.chan_list[cta_first];
        const ChanParams &p = a.params[ch];
        ChanState *st = a.state + ch;
        const unsigned gmask = 0x11111111u << cl;
        // ring rows of this channel: I in column group g / 8, Q in column group 4 + g / 8
        unsigned char *xi1 = sm.xring[0] + (g >> 3) * XSBO + (g & 7) * 16, *xq1 = xi1 + 4 * XSBO;
        unsigned char *xi2 = sm.xring[1] + (g >> 3) * XSBO + (g & 7) * 16, *xq2 = xi2 + 4 * XSBO;

        FirLaneState ls;
        ls.te1 = st->teta1_old; ls.te2 = st->teta2_old; ls.te3 = st->teta3_old; ls.c1 = st->M_c1; ls.c2 = st->M_c2; ls.clip = 0;
        const int iq_auto = p.iq_auto, shift_kind = p.shift_kind, shift_down = p.shift_down;
        const bool any_auto = __any_sync(0xffffffffu, iq_auto != 0);
        const bool fast_fe = __all_sync(0xffffffffu, iq_auto != 0 && shift_kind == 1);
        const float adj_i = p.adj_i, adj_q = p.adj_q, phase_bal = p.phase_bal;
        const size_t chan_base = (size_t)ch * (size_t)a.chan_stride;
        // the lane's first sample: block hb of pass 0, sample 4 q of the block
        const int2 *__restrict__ src = reinterpret_cast<const int2 *>(a.iq) + chan_base + 32 * hb + 4 * (q_lo + 2 * q_hi);
        auto ld256 = [](int4 &lo, int4 &hi, const int2 *ptr) {
            asm volatile("ld.global.nc.v8.s32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(ptr));
        };
        // input prefetch: 2 x 2 x 32 bytes = the 8 samples of either pass of the next step, straight from global memory into registers
        // (pre[2 j], pre[2 j + 1] = samples 16 j + 4 q .. + 3 of the block)
        int4 pre[4], pre1[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { pre[i] = make_int4(0, 0, 0, 0); pre1[i] = make_int4(0, 0, 0, 0); }
        if (active) {
#pragma unroll
            for (int j = 0; j < 2; j++) { ld256(pre[2 * j], pre[2 * j + 1], src + 16 * j); ld256(pre1[2 * j], pre1[2 * j + 1], src + 64 + 16 * j); }
        }

        PROF_DECL;
        for (int it = 0; it < niter; it++) {
            PROF_TOP(it);
            if (it == V - 1 && active) {
                // decimator history x[-96..-1] (82 used) -> the last 96 slots of virtual step V - 1
                for (int b = rp; b < 96; b += 8) {
                    unsigned i1, i2, q1, q2;
                    split_bf16(st->s1_hist_i[b], i1, i2); split_bf16(st->s1_hist_q[b], q1, q2);
                    const int off = (((V - 1) & 1) * CH4 + 32 + b);
                    const int bo = (off >> 3) * XLBO + (off & 7) * 2;
                    *reinterpret_cast<unsigned short *>(xi1 + bo) = (unsigned short)i1; *reinterpret_cast<unsigned short *>(xi2 + bo) = (unsigned short)i2;
                    *reinterpret_cast<unsigned short *>(xq1 + bo) = (unsigned short)q1; *reinterpret_cast<unsigned short *>(xq2 + bo) = (unsigned short)q2;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            if (it >= V && it < s_end && !KNOCK(1)) {
                const int t = it - V;
                // The MMAs of step it - 2 read the ring buffer written now; they were committed one iteration ago.
                mbar_wait(&sm.bar_dec[it & 1], (unsigned)(((it - 2) >> 1) & 1));
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    // ---- front end of one pass.  The 2^-16 input scaling (audio_driver.c:2680-2685) is exact, so it is folded
                    // into the correction factors; the Fs/4 translation (freq_shift.c:219-262) is a sign/swap pattern of
                    // period 4 folded into the same factors.
                    float fi[8], fq[8];
                    int lvmax = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int4 v = pre[i];
                        lvmax = max(lvmax, max(abs(v.x), abs(v.z)));
                        fi[2 * i] = (float)v.x; fq[2 * i] = (float)v.y; fi[2 * i + 1] = (float)v.z; fq[2 * i + 1] = (float)v.w;
                    }
                    if (h == 0) {
#pragma unroll
                        for (int i = 0; i < 4; i++) pre[i] = pre1[i];      // the second pass runs on the other half of the prefetch
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
                        for (int k = 0; k < 8; k++) {
                            s1 += __uint_as_float(__float_as_uint(fq[k]) ^ (__float_as_uint(fi[k]) & 0x80000000u));
                            s2 += fabsf(fi[k]); s3 += fabsf(fq[k]);
                        }
                        // the four lanes of a block: q_lo, q_hi = lane bits 2, 4
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 4); s2 += __shfl_xor_sync(0xffffffffu, s2, 4); s3 += __shfl_xor_sync(0xffffffffu, s3, 4);
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 16); s2 += __shfl_xor_sync(0xffffffffu, s2, 16); s3 += __shfl_xor_sync(0xffffffffu, s3, 16);
                        // first-order low-pass over the two blocks of the pass (:2281-2283), then M_c1 / M_c2 (:2285-2295) of the own block
                        float t1 = ls.te1, t2 = ls.te2, t3 = ls.te3, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
                        const float kE = 0.003f * 0.03125f * kS;
                        float bs1[2], bs2[2], bs3[2];
#pragma unroll
                        for (int b = 0; b < 2; b++) {
                            bs1[b] = __shfl_sync(0xffffffffu, s1, 8 * b + cl); bs2[b] = __shfl_sync(0xffffffffu, s2, 8 * b + cl); bs3[b] = __shfl_sync(0xffffffffu, s3, 8 * b + cl);
                        }
#pragma unroll
                        for (int b = 0; b < 2; b++) {
                            t1 = fmaf(0.997f, t1, -kE * bs1[b]); t2 = fmaf(0.997f, t2, kE * bs2[b]); t3 = fmaf(0.997f, t3, kE * bs3[b]);
                            if (hb == b) { m1 = t1; m2 = t2; m3 = t3; }
                        }
                        const float den = m2 * m2;
                        const float hlp = (den > 0.0f) ? __fdividef(fmaf(m3, m3, -m1 * m1), den) : den;
                        if (iq_auto) {
                            ls.te1 = t1; ls.te2 = t2; ls.te3 = t3;
                            c1m = (m2 != 0.0f) ? __fdividef(m1, m2) : 0.0f;
                            c2m = (hlp > 0.0f) ? hlp * rsqrtf(hlp) : 1.0f;
                            ls.c1 = c1m; ls.c2 = c2m;                            // after the second pass, the lanes of block 1 hold the block-3 values the state keeps
                        }
                    }
                    if (a.iqc_log != nullptr && q_lo + q_hi == 0 && active) {
                        // the spectrum tap kernel re-applies the correction to the last 16 blocks of the call: log their factors
                        const int j = 4 * t + 2 * h + hb - (a.nblocks - min(16, a.nblocks));
                        if (j >= 0) a.iqc_log[(size_t)(cta_first + g) * 16 + j] = make_float2(c1m, c2m);
                    }
                    if (fast_fe) {
                        // every channel of the warp: automatic IQ correction + Fs/4 translation (the default).
                        //   i' = c2 i, q' = q + c1 i;  phase 0: (i', q')  1: (q', -i')  2: (-i', -q')  3: (-q', i'), (x sgd when translating down)
                        const float sgd = shift_down ? -1.0f : 1.0f;
                        const float fa = c2m * kS, fd = c1m * kS, fas = fa * sgd, fds = fd * sgd, ks = kS * sgd;
#pragma unroll
                        for (int k = 0; k < 8; k += 4) {
                            const float i0 = fi[k], i1 = fi[k + 1], i2 = fi[k + 2], i3 = fi[k + 3];
                            fi[k] = i0 * fa;                                  fq[k] = fmaf(i0, fd, fq[k] * kS);
                            fi[k + 1] = fmaf(i1, fds, fq[k + 1] * ks);        fq[k + 1] = i1 * -fas;
                            fi[k + 2] = i2 * -fa;                             fq[k + 2] = fmaf(i2, -fd, fq[k + 2] * -kS);
                            fi[k + 3] = fmaf(i3, -fds, fq[k + 3] * -ks);      fq[k + 3] = i3 * fas;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; k++) {
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
                            fi[k] = vi; fq[k] = vq;
                        }
                    }
                    // ---- bf16 split, decimator ring: samples 16 j + 4 q .. + 3 of block hb of pass h of buffer it & 1 = half a 16-byte
                    // row (time group 8 h + 4 hb + 2 j + q_hi, bytes 8 q_lo .. + 7) per array and j
                    {
                        unsigned wi1[4], wi2[4], wq1[4], wq2[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            split_pack2(fi[2 * k], fi[2 * k + 1], wi1[k], wi2[k]);
                            split_pack2(fq[2 * k], fq[2 * k + 1], wq1[k], wq2[k]);
                        }
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int bo = ((it & 1) * (CH4 / 8) + 8 * h + 4 * hb + 2 * j + q_hi) * XLBO + 8 * q_lo;
                            *reinterpret_cast<uint2 *>(xi1 + bo) = make_uint2(wi1[2 * j], wi1[2 * j + 1]);
                            *reinterpret_cast<uint2 *>(xi2 + bo) = make_uint2(wi2[2 * j], wi2[2 * j + 1]);
                            *reinterpret_cast<uint2 *>(xq1 + bo) = make_uint2(wq1[2 * j], wq1[2 * j + 1]);
                            *reinterpret_cast<uint2 *>(xq2 + bo) = make_uint2(wq2[2 * j], wq2[2 * j + 1]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // visible to the tensor core after the step barrier
                // fetch the next step; issued after the fence (a membar that would wait for these loads) and
                // consumed after the step barrier, where this warp waits for the slower roles anyway (both passes: a load issued
                // inside the step does not arrive in time for its second pass)
                if (t + 1 < nsteps && active) {
                    const int2 *nx = src + (size_t)(t + 1) * CH4;
#pragma unroll
                    for (int j = 0; j < 2; j++) { ld256(pre[2 * j], pre[2 * j + 1], nx + 16 * j); ld256(pre1[2 * j], pre1[2 * j + 1], nx + 64 + 16 * j); }
                }
            }
            PROF_END(it);
            __syncthreads();
        }
        PROF_SAVE();
        // ---- store state: decimator history = the last 96 samples of the last step, IQ-correction state ----
        if (active) {
            const int sl = (s_end - 1) & 1;
            for (int b = rp; b < 96; b += 8) {
                const int off = sl * CH4 + 32 + b;
                const int bo = (off >> 3) * XLBO + (off & 7) * 2;
                const unsigned i1 = *reinterpret_cast<const unsigned short *>(xi1 + bo), i2 = *reinterpret_cast<const unsigned short *>(xi2 + bo);
                const unsigned q1 = *reinterpret_cast<const unsigned short *>(xq1 + bo), q2 = *reinterpret_cast<const unsigned short *>(xq2 + bo);
                st->s1_hist_i[b] = join_bf16(i1, i2);
                st->s1_hist_q[b] = join_bf16(q1, q2);
            }
            int clip = ls.clip;
            ls.c1 = __shfl_sync(gmask, ls.c1, 8 + cl); ls.c2 = __shfl_sync(gmask, ls.c2, 8 + cl);
            clip |= __shfl_xor_sync(gmask, clip, 4); clip |= __shfl_xor_sync(gmask, clip, 8); clip |= __shfl_xor_sync(gmask, clip, 16);
            if (rp == 0) {
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
        const unsigned idesc_d = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(64 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const unsigned idesc_h = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(32 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const unsigned idesc_d2 = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(128 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const unsigned idesc_h2 = idesc_d;
        const unsigned long long ad1 = umma_desc(smem_u32(sm.gd[0]) + DR0 * 16, DPLANE, 128), ad2 = umma_desc(smem_u32(sm.gd[1]) + DR0 * 16, DPLANE, 128);
        static_assert(sizeof(sm.xring[0]) == 8 * XSBO && sizeof(sm.hring[0]) == 4 * HSBO, "the lo half of a ring is the next column groups of the hi half");
        const unsigned long long bx1 = umma_desc(smem_u32(sm.xring[0]), XLBO, XSBO);
        unsigned long long ah[4], bh[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { ah[i] = umma_desc(smem_u32(sm.gh[i]) + HR0 * 16, HPLANE, 128); bh[i] = umma_desc(smem_u32(sm.hring[i]), HLBO, HSBO); }
        const int oc_last = (s_end - 1) >> 2;
        // The operand fetch of the MMAs saturates shared memory; issued back to back they starve the LDS of the serial warps for
        // the length of the burst.  A short pause after every k-step leaves gaps for them (the MMAs have the whole iteration).
#ifdef UHSDR_TC_PROF
        const int mma_pause = g_tc_pause >= 0 ? g_tc_pause : MMA_PAUSE;
#else
        constexpr int mma_pause = MMA_PAUSE;
#endif
        auto pause = [&]() { if (mma_pause > 0) { const long long t0 = clock64(); while (clock64() - t0 < mma_pause) { } } };
        PROF_DECL;
        for (int it = 0; it < niter; it++) {
            PROF_TOP(it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const int sd = it - 1;               // step whose samples were written into the decimator ring in the previous iteration
                if (sd >= 0) {
                    const unsigned long long b1 = bx1 + (unsigned)((sd & 1) * (CH4 / 8) * (XLBO / 16));
                    if (sd >= V && sd < s_end) {
                        // the 8 k-steps of this step into its own chunk: kk = 6 + 8 (sd & 3) + j
                        const unsigned d_tmem = tmem + DEC_COL0 + (unsigned)(((sd >> 2) & 1) * DEC_COLS);
                        const unsigned arow = (unsigned)(4 * (6 + 8 * (sd & 3)));
#pragma unroll 1
                        for (int j = 0; j < 8; j++) {
                            if (KNOCK(2)) continue;
                            umma2_bf16(d_tmem, ad1 - arow - 4 * j, ad2 - arow - 4 * j, b1 + (2 * XLBO / 16) * j, idesc_d2, idesc_d, 1u);
                            pause();
                        }
                    }
                    if ((sd & 3) == 3 && sd >= V - 1 && sd + 1 < s_end) {
                        // the last 96 samples of the step are the history of the next chunk: kk = 0 .. 5, first write of its accumulator
                        const unsigned d_tmem = tmem + DEC_COL0 + (unsigned)((((sd >> 2) + 1) & 1) * DEC_COLS);
#pragma unroll 1
                        for (int j = 2; j < 8; j++) {
                            if (KNOCK(2)) continue;
                            umma2_bf16(d_tmem, ad1 - 4 * (j - 2), ad2 - 4 * (j - 2), b1 + (2 * XLBO / 16) * j, idesc_d2, idesc_d, j > 2 ? 1u : 0u);
                            pause();
                        }
                    }
                    umma_commit(&sm.bar_dec[sd & 1]);
                }
                const int sh = it - 3;               // step whose decimator outputs were written into the Hilbert ring in the previous iteration
                if (sh >= 0) {
                    if (sh >= 1 && sh < s_end) {
#pragma unroll 1
                        for (int e = 0; e < 2; e++) {
                            const int u = 2 * sh + e;                        // global k-step (16 decimated samples)
                            const int oa = u >> 3, ob = u & 7;
                            const unsigned boff = (unsigned)(((sh & 1) * (ND / 8) + 2 * e) * (HLBO / 16));
#pragma unroll 1
                            for (int dd = 0; dd < 3; dd++) {
                                const int oc = oa + dd;                      // output chunk this slab contributes to
                                if ((dd < 2 || ob >= 3) && oc >= V / 4 && oc <= oc_last && !KNOCK(2)) {
                                    const int kk = u - 8 * oc + 13;
                                    const unsigned d_tmem = tmem + HIL_COL0 + (unsigned)((oc & 3) * HIL_COLS);
                                    umma2_bf16(d_tmem, ah[0] - 16 * kk, ah[1] - 16 * kk, bh[0] + boff, idesc_h2, idesc_h, kk > 0 ? 1u : 0u);     // [I1 | I2]
                                    umma2_bf16(d_tmem, ah[2] - 16 * kk, ah[3] - 16 * kk, bh[2] + boff, idesc_h2, idesc_h, 1u);                    // [Q1 | Q2]
                                    pause();
                                }
                            }
                        }
                    }
                    umma_commit(&sm.bar_hil[sh & 1]);
                }
            }
            __syncwarp();
            PROF_END(it);
            __syncthreads();
        }
        PROF_SAVE();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
        return;
    }

    if (warp >= W_EPI && warp < W_EPI + 4) {
        // ======================= epilogue warps: TMEM lane quadrant warp % 4 =====================
        // Step s occupies rows 32 (s & 3) .. +31 of its chunk accumulator = the lanes of quadrant s & 3; lane = decimated sample.
        const int qd = warp & 3;
        const unsigned lane_base = (unsigned)(qd * 32) << 16;
        unsigned lsbmask;
        float g_inv_max, g_slope, g_target, g_fixed;
        bool g_off;
        {
            const int chn = sm.chan[lane];
            lsbmask = __ballot_sync(0xffffffffu, lane < n_here && a.params[chn].lsb != 0);
            const AgcP &ap = a.params[chn].agc;
            g_inv_max = ap.inv_max_input; g_slope = ap.slope_constant; g_target = ap.out_target; g_fixed = ap.fixed_gain; g_off = ap.mode == 5;
        }
        const int gcol = lane < n_here ? lane : 0;
        PROF_DECL;
        for (int it = 0; it < niter; it++) {
            PROF_TOP(it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int sd = it - 2;
            if (sd >= 1 && sd < s_end && (sd & 3) == qd) {
                const int slot = (sd & 1) * ND + lane;
                const int bo = (slot >> 3) * HLBO + (slot & 7) * 2;
                unsigned char *h0 = sm.hring[0] + bo, *h1 = sm.hring[1] + bo, *h2 = sm.hring[2] + bo, *h3 = sm.hring[3] + bo;
                // state slot of this lane's sample once the launch is over (s2_hist[2..199] = the newest 198 decimator outputs)
                const int i_new = ND * (sd - V) + lane - ND * nsteps + 200;
                if (sd < V) {
                    // virtual step: Hilbert history d[-198..-1] = state slots 2..199 (already rounded if this kernel wrote them)
                    const int i_old = ND * (sd - V) + lane + 200;
                    if (sd >= 2) mbar_wait(&sm.bar_hil[sd & 1], (unsigned)(((sd - 2) >> 1) & 1));
#pragma unroll 1
                    for (int n = 0; n < FG; n++) {
                        float vi = 0.0f, vq = 0.0f;
                        if (n < n_here && i_old >= 2) {
                            ChanState *stn = a.state + sm.chan[n];
                            vi = stn->s2_hist_i[i_old]; vq = stn->s2_hist_q[i_old];
                        }
                        unsigned i1, i2, q1, q2;
                        split_bf16(vi, i1, i2); split_bf16(vq, q1, q2);
                        if (n < n_here && i_old >= 2 && i_new >= 2) { a.state[sm.chan[n]].s2_hist_i[i_new] = join_bf16(i1, i2); a.state[sm.chan[n]].s2_hist_q[i_new] = join_bf16(q1, q2); }
                        if ((lsbmask >> n) & 1u) { q1 ^= 0x8000u; q2 ^= 0x8000u; }
                        const int co = (n >> 3) * HSBO + (n & 7) * 16;
                        *reinterpret_cast<unsigned short *>(h0 + co) = (unsigned short)i1; *reinterpret_cast<unsigned short *>(h1 + co) = (unsigned short)i2;
                        *reinterpret_cast<unsigned short *>(h2 + co) = (unsigned short)q1; *reinterpret_cast<unsigned short *>(h3 + co) = (unsigned short)q2;
                    }
                } else {
                    mbar_wait(&sm.bar_dec[sd & 1], (unsigned)((sd >> 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned taddr = tmem + DEC_COL0 + (unsigned)(((sd >> 2) & 1) * DEC_COLS) + lane_base;
                    // the Hilbert MMAs of step sd - 2 read this ring buffer; they were committed one iteration ago
                    mbar_wait(&sm.bar_hil[sd & 1], (unsigned)(((sd - 2) >> 1) & 1));
                    const bool save = i_new >= 2;
#pragma unroll 1
                    for (int gq = 0; gq < (KNOCK(4) ? 0 : FG / 4); gq++) {         // 4 channels per pass (rolled: the instruction cache is the scarce resource)
                        unsigned vi[4], vq[4], wi[4], wq[4];
                        TMEM_LD_X4(vi, taddr + (unsigned)(4 * gq));
                        TMEM_LD_X4(vq, taddr + (unsigned)(32 + 4 * gq));
                        TMEM_LD_X4(wi, taddr + (unsigned)(64 + 4 * gq));            // the c1 x2 halves
                        TMEM_LD_X4(wq, taddr + (unsigned)(96 + 4 * gq));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int n8 = 0; n8 < 4; n8++) {
                            vi[n8] = __float_as_uint(__fadd_rn(__uint_as_float(vi[n8]), __uint_as_float(wi[n8])));
                            vq[n8] = __float_as_uint(__fadd_rn(__uint_as_float(vq[n8]), __uint_as_float(wq[n8])));
                        }
                        const unsigned lm = lsbmask >> (4 * gq);
                        const int cbase = (gq >> 1) * HSBO + (gq & 1) * 64;
#pragma unroll
                        for (int n8 = 0; n8 < 4; n8++) {
                            unsigned i1, i2, q1, q2;
                            split_bf16(__uint_as_float(vi[n8]), i1, i2); split_bf16(__uint_as_float(vq[n8]), q1, q2);
                            if (save && 4 * gq + n8 < n_here) {
                                ChanState *stn = a.state + sm.chan[4 * gq + n8];
                                stn->s2_hist_i[i_new] = join_bf16(i1, i2); stn->s2_hist_q[i_new] = join_bf16(q1, q2);
                            }
                            if ((lm >> n8) & 1u) { q1 ^= 0x8000u; q2 ^= 0x8000u; }       // LSB: I - Q
                            const int co = cbase + n8 * 16;
                            *reinterpret_cast<unsigned short *>(h0 + co) = (unsigned short)i1; *reinterpret_cast<unsigned short *>(h1 + co) = (unsigned short)i2;
                            *reinterpret_cast<unsigned short *>(h2 + co) = (unsigned short)q1; *reinterpret_cast<unsigned short *>(h3 + co) = (unsigned short)q2;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // visible to the tensor core after the step barrier
            }
            const int sh = it - 4;
            if (sh >= V && sh < s_end && (sh & 3) == qd && !KNOCK(8)) {
                // ---- Hilbert outputs of step sh leave TMEM: lane = decimated sample, column = channel ----
                mbar_wait(&sm.bar_hil[sh & 1], (unsigned)((sh >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned taddr = tmem + HIL_COL0 + (unsigned)(((sh >> 2) & 3) * HIL_COLS) + lane_base;
                float *aud = sm.aud[sh & 1] + lane * SMS;
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += 16) {                                 // two halves of 16 channels: c1 x1 + c2 x1 | c1 x2
                    unsigned v[16], w[16];
                    TMEM_LD_X16(v, taddr + (unsigned)c0);
                    TMEM_LD_X16(w, taddr + (unsigned)(32 + c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; c++)
                        if (c0 + c < FG) aud[c0 + c] = __fadd_rn(__uint_as_float(v[c]), __uint_as_float(w[c]));
                }
            }
            // ---- AGC gain law (audio_agc.c:563-570; mode 5 = fixed gain :354-365) for the step the detector finished in the previous
            // iteration: elementwise, lane = channel, on the two warps that have no epilogue in this iteration (16 samples each).
            // out = x[n - 49] * (out_target - slope_constant * min(0, log10f_fast(volts / max_input))) / volts
            {
                const int c = it - IT_GAIN;
                const int k1 = (qd - it - 1) & 3;                       // 0: first half, 2: second half, 1 / 3: busy with an epilogue
                if (c >= 0 && c < nsteps && !(k1 & 1) && !KNOCK(16)) {
                    const int i0 = k1 * 8;
                    const float *vin = sm.agc[c & 1] + gcol;
                    float *out = sm.gq[c & 1] + gcol;
                    const float *latp = sm.lat + gcol;
                    int ra = (c % 5) * ND + i0 - (g_off ? 0 : AGC_W); if (ra < 0) ra += LR;
#pragma unroll 1
                    for (int i8 = 0; i8 < 16; i8 += 4) {
                        float xv[4], vv[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            int rr = ra + i8 + i; if (rr >= LR) rr -= LR;
                            vv[i] = vin[(i0 + i8 + i) * SMS]; xv[i] = latp[rr * SMS];
                        }
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            // Math_log10f_fast (uhsdr_math.c:27-41) of inv_max_input * volts (> 0): exponent / mantissa by bit
                            // operations, the cubic in Horner form
                            const unsigned ub = __float_as_uint(__fmul_rn(g_inv_max, vv[i]));
                            const float F = __uint_as_float((ub & 0x007fffffu) | 0x3f000000u);      // frexpf mantissa in [0.5, 1)
                            const float E = (float)((int)(ub >> 23) - 126);
                            float Y = fmaf(1.23149591368684f, F, -4.11852516267426f);
                            Y = fmaf(Y, F, 6.02197014179219f);
                            Y = fmaf(Y, F, -3.13396450166353f);
                            float vo = __fmul_rn(__fadd_rn(Y, E), 0.3010299956639812f);
                            vo = fminf(vo, 0.0f);
                            const float mult = g_off ? g_fixed : __fdividef(fmaf(-g_slope, vo, g_target), vv[i]);
                            xv[i] = __fmul_rn(xv[i], mult);
                        }
                        if (lane < n_here) {
#pragma unroll
                            for (int i = 0; i < 4; i++) out[(i0 + i8 + i) * SMS] = xv[i];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            PROF_END(it);
            __syncthreads();
        }
        PROF_SAVE();
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
        PROF_DECL;
        for (int t = 0; t < niter; t++) {
            PROF_TOP(t);
            const int c = t - IT_LAT;
            if (c >= 0 && c < nsteps && !KNOCK(32)) {
                const float *in = sm.aud[c & 1];
                float *out = sm.lat + (c % 5) * ND * SMS;
                // Two samples per loop iteration: the loop body (~80 instructions) then stays resident in the 6 KB level-0
                // instruction cache this warp shares with three warps that run other code; longer bodies are refetched line
                // by line from the next cache level at ~50 cycles per 8 instructions.  The inputs of the next iteration are
                // fetched before the current one is computed, so their shared-memory latency stays off the recurrence.
                constexpr int LG = 2;
                float nxt[2][LG];                        // two iterations ahead
                const float *pin = in + gq;
                float *pout = out + gq;
#pragma unroll
                for (int i = 0; i < LG; i++) { nxt[0][i] = pin[i * SMS]; nxt[1][i] = pin[(LG + i) * SMS]; }
#pragma unroll 1
                for (int i0 = 0; i0 < ND; i0 += LG) {
                    float xin[LG], yo[LG];
                    const int inx = min(i0 + 2 * LG, ND - LG);
#pragma unroll
                    for (int i = 0; i < LG; i++) { xin[i] = nxt[0][i]; nxt[0][i] = nxt[1][i]; nxt[1][i] = pin[(inx + i) * SMS]; }
#pragma unroll
                    for (int i = 0; i < LG; i++) {
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
                        for (int i = 0; i < LG; i++) pout[(i0 + i) * SMS] = yo[i];
                    }
                }
            }
            PROF_END(t);
            __syncthreads();
        }
        PROF_SAVE();
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
        // The suffix maxima of step c live in buffer c % 3; the biquad warp, which has time to spare, computes them (and, when no
        // channel uses the hang AGC, the hang back-average, which then feeds nothing but the status word) while this warp works
        // on step c with the buffers of steps c - 1 and c - 2.  Steps -1 and -2 (the history) are prepared here.
        if (active) {
            float m = 0.0f;
            for (int o = ND - 1; o >= 0; o--) {
                m = fmaxf(m, fabsf(latp[(LR - ND + o) * SMS]));
                sm.smax[2][o * SMS + g] = m;
            }
            m = 0.0f;
            for (int o = ND - 1; o >= 16; o--) {
                m = fmaxf(m, fabsf(latp[(LR - 2 * ND + o) * SMS]));
                sm.smax[1][o * SMS + g] = m;
            }
        }
        __syncwarp();
        const bool any_hang = __any_sync(0xffffffffu, active && (ap.hang_enable || ar.state == 2 || ar.state == 4 || ar.decay_type != 0 || ar.hang_counter > 0));
        PROF_DECL;
        for (int t = 0; t < niter; t++) {
            PROF_TOP(t);
            const int c = t - IT_AGC;
            if (c >= 0 && c < nsteps && active && ap.mode != 5 && !KNOCK(64)) {
                const int row0 = (c % 5) * ND;                  // ring row of the first sample of this step
                const float *in = latp + row0 * SMS;
                float *out = sm.agc[c & 1] + g;
                const float *S1 = sm.smax[(c + 2) % 3] + g, *S2 = sm.smax[(c + 1) % 3] + g;      // steps c - 1, c - 2
                const float mprev = S1[0];
                float pmax = 0.0f;
                auto detect = [&](auto hangc) {
                    constexpr bool HANG = decltype(hangc)::value;
                // operands of one group of AG samples.  The delayed sample x[n-49] of group element j sits at ring row
                // (row0 + k8 + j - 49) mod LR; a group wraps at most between its first and second element.
                // Operands of the next group of AG = 2 samples, fetched with running pointers (no per-group index arithmetic).
                // The delayed sample x[n-49] of element j sits at ring row (row0 + k8 + j - 49) mod LR; the sliding-maximum
                // operand comes from the suffix maxima of the step before the previous one while k8 < 16 (combined with the
                // whole previous step, mprev), afterwards from those of the previous step (mp = 0: all values are >= 0).
                // The last two calls run past the step; what they fetch is valid shared memory and never used.
                int ra0 = row0 - AGC_W; if (ra0 < 0) ra0 += LR;
                const float *const lat_end = latp + LR * SMS;
                const float *px = in, *pa = latp + ra0 * SMS, *pb = pa + SMS, *pc = S2 + 16 * SMS;
                if (pb >= lat_end) pb -= LR * SMS;
                float mp = mprev;
                int kl = 0;
                float *po = out;
                auto load_next = [&](float (&x)[AG], float (&dly)[AG], float (&cmx)[AG]) {
                    static_assert(AG == 2, "two samples per group");
                    x[0] = px[0]; x[1] = px[SMS];
                    dly[0] = pa[0]; dly[1] = pb[0];
                    cmx[0] = fmaxf(mp, pc[0]); cmx[1] = fmaxf(mp, pc[SMS]);
                    px += AG * SMS; pc += AG * SMS; kl += AG;
                    pa += AG * SMS; if (pa >= lat_end) pa -= LR * SMS;
                    pb += AG * SMS; if (pb >= lat_end) pb -= LR * SMS;
                    if (kl == 16) { pc = S1; mp = 0.0f; }
                };
                // the operands of the next two iterations are fetched ahead of the current one (shared-memory latency off the recurrence)
                float xn[AG], dn[AG], cn[AG], xm[AG], dm[AG], cm[AG];
                load_next(xn, dn, cn);
                load_next(xm, dm, cm);
#pragma unroll 1
                for (int k8 = 0; k8 < ND; k8 += AG) {
                    float x[AG], dly[AG], cmx[AG];
#pragma unroll
                    for (int j = 0; j < AG; j++) { x[j] = xn[j]; dly[j] = dn[j]; cmx[j] = cn[j]; xn[j] = xm[j]; dn[j] = dm[j]; cn[j] = cm[j]; }
                    load_next(xm, dm, cm);
#pragma unroll
                    for (int j = 0; j < AG; j++) {
                        const float abs_out = fabsf(dly[j]), abs_in = fabsf(x[j]);
                        pmax = fmaxf(pmax, abs_in);
                        ar.fast_backaverage = fmaf(ap.fast_backmult, abs_out, __fmul_rn(ap.onemfast_backmult, ar.fast_backaverage));
                        if constexpr (HANG) ar.hang_backaverage = fmaf(ap.hang_backmult, abs_out, __fmul_rn(ap.onemhang_backmult, ar.hang_backaverage));
                        ar.ring_max = fmaxf(pmax, cmx[j]);
                        const float dv = __fsub_rn(ar.ring_max, ar.volts);
                        const bool attack = ar.ring_max >= ar.volts;
                        float mult_sel = ap.attack_mult;
                        bool upd = true;
                        int nstate = ar.state;
                        if constexpr (!HANG) {
                            // hang AGC disabled on every channel of this warp (the default, ui_configuration.c:81): only
                            // states 0 / 1 / 3 occur and the 5-state machine reduces to selects
                            // (bitwise on purpose: short-circuit evaluation turns into divergent branches in this serial loop)
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
                        if (upd) ar.volts = fmaf(dv, mult_sel, ar.volts);
                        if (ar.volts < ap.min_volts) { ar.volts = ap.min_volts; ar.action = 0; } else { ar.action = 1; }
                        po[j * SMS] = ar.volts;
                    }
                    po += AG * SMS;
                }
                };
                if (any_hang) detect(std::true_type{}); else detect(std::false_type{});
            }
            PROF_END(t);
            __syncthreads();
        }
        PROF_SAVE();
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
            st->agc_fast_backaverage = ar.fast_backaverage;
            st->agc_hang_counter = ar.hang_counter; st->agc_decay_type = ar.decay_type; st->agc_state = ar.state;
            st->agc_action = ar.action;
            if (any_hang) {          // otherwise the biquad warp keeps the hang back-average
                st->agc_hang_backaverage = ar.hang_backaverage;
                st->agc_hang_action = (ar.hang_backaverage > ap.hang_level) ? 1 : 0;
            }
        }
        return;
    }

    if (warp == W_BQ) {
        // ---- the fixed gain (:2513-2524) and the 4-stage DF1 cascade IIR_biquad_1 (:2527) at 12 ksps.  Per stage the terms
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
        float xl1 = 0.0f, xl2 = 0.0f;        // the last two cascade inputs (state of skipped leading stages)
        float yl[3];                         // the last three cascade outputs: the interpolator's history (arm_fir_interpolate_f32 state)
#pragma unroll
        for (int q = 0; q < 3; q++) yl[q] = st->interp_hist[INTERP_HIST - 3 + q];
        // one step of 32 samples with the skipped stages known at compile time
        auto run_step = [&](auto maskc, const float *in, float *out) {
            constexpr unsigned MASK = decltype(maskc)::value;
#pragma unroll 1
            for (int i0 = 0; i0 < ND; i0 += 8) {
                float xv[8];
#pragma unroll
                for (int i = 0; i < 8; i++) xv[i] = in[(i0 + i) * SMS];
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
                yl[0] = xv[5]; yl[1] = xv[6]; yl[2] = xv[7];
            }
        };
        if (skipmask != 0xbu) skipmask = 0;      // only the default plan (bass shelf alone) has a specialised loop
        // Helper for the AGC detector (the longest of the serial roles; this warp is the shortest): while W_AGC works on step c,
        // this warp computes the suffix maxima of |x| of step c (needed from step c + 1 on) and, when no channel of the CTA
        // uses the hang AGC, the hang back-average over x[n - 49] (audio_agc.c:400-407).
        const AgcP &hp = p.agc;
        const bool agc_on = active && hp.mode != 5;
        const bool any_hang = __any_sync(0xffffffffu, active && (hp.hang_enable || st->agc_state == 2 || st->agc_state == 4 || st->agc_decay_type != 0 ||
                                                                  st->agc_hang_counter > 0));
        const float hbm = hp.hang_backmult, ohbm = hp.onemhang_backmult;
        float hba = st->agc_hang_backaverage;
        PROF_DECL;
        for (int t = 0; t < niter; t++) {
            PROF_TOP(t);
            const int ca = t - IT_AGC;
            if (ca >= 0 && ca < nsteps && agc_on && !KNOCK(64)) {
                const float *latp = sm.lat + g;
                const float *in = latp + (ca % 5) * ND * SMS;
                float *Sn = sm.smax[ca % 3] + g;
                int ra = (ca % 5) * ND - AGC_W; if (ra < 0) ra += LR;
                float m = 0.0f;
#pragma unroll 1
                for (int o8 = ND - 8; o8 >= 0; o8 -= 8) {
                    float rv[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) rv[j] = in[(o8 + j) * SMS];
#pragma unroll
                    for (int j = 7; j >= 0; j--) { m = fmaxf(m, fabsf(rv[j])); Sn[(o8 + j) * SMS] = m; }
                }
                if (!any_hang) {
#pragma unroll 1
                    for (int o8 = 0; o8 < ND; o8 += 8) {
                        float dl[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) { int rr = ra + o8 + j; if (rr >= LR) rr -= LR; dl[j] = latp[rr * SMS]; }
#pragma unroll
                        for (int j = 0; j < 8; j++) hba = fmaf(hbm, fabsf(dl[j]), __fmul_rn(ohbm, hba));
                    }
                }
            }
            const int c = t - IT_BQ;
            if (c >= 0 && c < nsteps && active && a.nr_handoff) {
                // spectral NR follows: the AGC output of this step leaves for the scratch row of the channel (8 floats per block)
                const float *in = sm.gq[c & 1] + g;
                float4 *dst = reinterpret_cast<float4 *>(a.scratch + (size_t)(cta_first + g) * (size_t)a.scratch_stride + (size_t)c * ND);
#pragma unroll 1
                for (int i0 = 0; i0 < ND; i0 += 8) {
                    float xv[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) xv[i] = in[(i0 + i) * SMS];
                    dst[i0 / 4] = make_float4(xv[0], xv[1], xv[2], xv[3]); dst[i0 / 4 + 1] = make_float4(xv[4], xv[5], xv[6], xv[7]);
                }
            } else if (c >= 0 && c < nsteps && active && !KNOCK(128)) {
                const float *in = sm.gq[c & 1] + g;
                float *out = sm.bq[c & 1] + 3 * SMS + g;
#pragma unroll
                for (int q = 0; q < 3; q++) out[(q - 3) * SMS] = yl[q];
                if (skipmask == 0xbu) run_step(std::integral_constant<unsigned, 0xbu>{}, in, out);
                else run_step(std::integral_constant<unsigned, 0u>{}, in, out);
            }
            PROF_END(t);
            __syncthreads();
        }
        PROF_SAVE();
        if (active && !a.nr_handoff) {
            // a skipped (pass-through) stage saw the output of the nearest computed stage before it
            float s1v = xl1, s2v = xl2;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if (skipmask & (1u << s)) { bs[s].x1 = s1v; bs[s].x2 = s2v; bs[s].y1 = s1v; bs[s].y2 = s2v; }
                else { s1v = bs[s].y1; s2v = bs[s].y2; }
                st->bq1[s] = bs[s];
            }
        }
        if (active && agc_on && !any_hang) {
            st->agc_hang_backaverage = hba;
            st->agc_hang_action = (hba > hp.hang_level) ? 1 : 0;
        }
        return;
    }
    if (a.nr_handoff) {
        // output warps: nothing to do (the serial kernel's phase 2 owns the stages behind the NR); keep the step barriers
        for (int t = 0; t < niter; t++) __syncthreads();
        return;
    }

    // W_POSTA / W_POSTB: x4 interpolator (:2560-2577), anti-alias lattice (:2581-2583, 6 stages, some paths), treble biquad
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
        const bool plain = (mute == nullptr) && (dst_f == nullptr);
        const bool tr_unity = __all_sync(0xffffffffu, !active || (tc[0] == 1.0f && tc[1] == -tc[3] && tc[2] == -tc[4] &&
                                                                  ts.x1 == ts.y1 && ts.x2 == ts.y2));
        // one 32-sample block (8 decimated samples -> 32 outputs = 256 bytes of the channel's row)
        // PLAIN: no mute array and no float copy of the audio asked for (the throughput case): no masking, no second store
        float nx[2][2];                       // the inputs of the next two loop iterations, fetched from the queue ahead of their use
        // the lean variant: no anti-alias lattice, unity treble, no mute array, no float copy (same arithmetic, fewer instructions)
        const bool lean = plain && !any_aa && tr_unity;
        // The lean variant has no recurrence left (the interpolator is a FIR), so its lanes need not be channels.  When every
        // channel of the CTA also uses the same interpolator (there is one per decimation factor among the paths without an
        // anti-alias filter), the two output warps turn the step around: LANE = DECIMATED SAMPLE, one channel after the other
        // (W_POSTA the even slots, W_POSTB the odd ones).  A warp store is then 1 KB of one channel's row -- eight full
        // 128-byte lines -- instead of 28 separate 32-byte sectors in 28 rows, which is what keeps the load/store data path,
        // the busiest unit of this kernel, free for the other roles.  Otherwise W_POSTA runs one channel per lane.
        const bool postB = warp == W_POSTB;
        const int ic0 = __shfl_sync(0xffffffffu, p.interp_c, 0), P0 = __shfl_sync(0xffffffffu, P, 0);
        const bool tlean = lean && __all_sync(0xffffffffu, !active || (p.interp_c == ic0 && P == P0));
        auto run_block = [&](auto aac, auto trc, auto plainc, const float *in, int blk, int4 *d4, float4 *df, bool muted) {
            constexpr bool AA = decltype(aac)::value, TR = decltype(trc)::value, PLAIN = decltype(plainc)::value;
            const int mm = (!PLAIN && muted) ? 0 : -1;    // external_mute: zeros out, all state advanced (:2845-2853)
#pragma unroll 1
            for (int h = 0; h < 4; h++) {               // two decimated samples per iteration (short body: level-0 instruction cache)
                float xv[2];
                const int nrow = min(8 * blk + 2 * h + 4, ND - 2);
#pragma unroll
                for (int i = 0; i < 2; i++) { xv[i] = nx[0][i]; nx[0][i] = nx[1][i]; nx[1][i] = in[(nrow + i) * SMS]; }
#pragma unroll
                for (int i = 0; i < 2; i++) {
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
                    const int pos = 2 * h + i;
                    const int w0 = format_audio_word(__fmul_rn(o[0], 10.0f)) & mm, w1 = format_audio_word(__fmul_rn(o[1], 10.0f)) & mm;   // LINE_OUT_SCALING_FACTOR (:2860)
                    const int w2 = format_audio_word(__fmul_rn(o[2], 10.0f)) & mm, w3 = format_audio_word(__fmul_rn(o[3], 10.0f)) & mm;
                    // the four output samples {l, r} x 4 = 32 bytes: one 256-bit store (STG.E.ENL2.256)
                    asm volatile("st.global.v8.b32 [%0], {%1, %1, %2, %2, %3, %3, %4, %4};" ::"l"(d4 + 2 * pos), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
                    if (!PLAIN && df) df[pos] = muted ? make_float4(0.0f, 0.0f, 0.0f, 0.0f)
                                            : make_float4(__fmul_rn(o[0], 10.0f), __fmul_rn(o[1], 10.0f), __fmul_rn(o[2], 10.0f), __fmul_rn(o[3], 10.0f));
                }
            }
        };
        PROF_DECL;
        for (int t = 0; t < niter; t++) {
            PROF_TOP(t);
            const int c = t - IT_POST;
            if (c >= 0 && c < nsteps && tlean && !KNOCK(256)) {
                // lane = decimated sample: rows lane .. lane + 3 of the queue are x[n-3] .. x[n]
                const float *col = sm.bq[c & 1] + lane * SMS;
                int2 *const obase = reinterpret_cast<int2 *>(a.audio) + ((size_t)c * CH4 + 4 * lane);
                const bool last = c == nsteps - 1;
                // operands of the next two channels are fetched before the current two are computed (shared-memory latency off the path)
                struct Ops { float x0, x1, x2, x3; int chn; };
                auto fetch = [&](int gi) {
                    Ops r;
                    const int gs = gi < n_here ? gi : 0;
                    r.x0 = col[gs]; r.x1 = col[SMS + gs]; r.x2 = col[2 * SMS + gs]; r.x3 = col[3 * SMS + gs];
                    r.chn = sm.chan[gs];
                    return r;
                };
                auto emit = [&](const Ops &v) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) o[j] = fmaf(v.x3, ic[j][3], fmaf(v.x2, ic[j][2], fmaf(v.x1, ic[j][1], __fmul_rn(v.x0, ic[j][0]))));
                    const int w0 = format_audio_word(__fmul_rn(o[0], 10.0f)), w1 = format_audio_word(__fmul_rn(o[1], 10.0f));   // LINE_OUT_SCALING_FACTOR (:2860)
                    const int w2 = format_audio_word(__fmul_rn(o[2], 10.0f)), w3 = format_audio_word(__fmul_rn(o[3], 10.0f));
                    int2 *d = obase + (size_t)v.chn * (size_t)a.chan_stride;
                    asm volatile("st.global.v8.b32 [%0], {%1, %1, %2, %2, %3, %3, %4, %4};" ::"l"(d), "r"(w0), "r"(w1), "r"(w2), "r"(w3));
                    if (last && lane >= ND - 3) {
                        // end of the launch: interpolator history = the last three samples; the (identity) treble stage's state follows the signal
                        ChanState *stn = a.state + v.chn;
                        stn->interp_hist[INTERP_HIST - 3 + (lane - (ND - 3))] = v.x3;
                        if (lane == ND - 1) {
                            for (int q = 0; q < INTERP_HIST - 3; q++) stn->interp_hist[q] = 0.0f;
                            BiquadS tsn; tsn.x1 = o[3]; tsn.y1 = o[3]; tsn.x2 = o[2]; tsn.y2 = o[2];
                            stn->bq2 = tsn;
                        }
                    }
                };
                // one channel per iteration (short body: the instruction cache is the scarce resource of this kernel)
                Ops nx = fetch(postB ? 1 : 0);
#pragma unroll 1
                for (int gi = postB ? 1 : 0; gi < n_here; gi += 2) {
                    const Ops v = nx;
                    nx = fetch(gi + 2);
                    emit(v);
                }
            } else if (c >= 0 && c < nsteps && active && !tlean && !KNOCK(256)) {
                const float *in = sm.bq[c & 1] + 3 * SMS + g;
                if (!postB) {
#pragma unroll
                    for (int i = 0; i < 2; i++) { nx[0][i] = in[i * SMS]; nx[1][i] = in[(2 + i) * SMS]; }
#pragma unroll 1
                    for (int blk = 0; blk < 4; blk++) {
                        int4 *d4 = dst + (size_t)c * 64 + blk * 16;
                        const bool muted = mute && mute[c * 4 + blk];
                        float4 *df = dst_f ? dst_f + (size_t)c * 32 + blk * 8 : nullptr;
                        if (any_aa) run_block(std::true_type{}, std::true_type{}, std::false_type{}, in, blk, d4, df, muted);
                        else if (tr_unity) run_block(std::false_type{}, std::false_type{}, std::false_type{}, in, blk, d4, df, muted);
                        else run_block(std::false_type{}, std::true_type{}, std::false_type{}, in, blk, d4, df, muted);
                    }
                }
            }
            PROF_END(t);
            __syncthreads();
        }
        PROF_SAVE();
        if (active && !tlean && !postB) {             // (the lane-per-sample variant has stored its state with the last step)
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

cudaError_t launch_rx_ssb_tc(const RxArgs &a, int dec_c, int hil_ci, int hil_cq, int sm_count, cudaStream_t stream)
{
    if (a.num_items <= 0 || a.nblocks <= 0) return cudaSuccess;
    if (a.nblocks % 4 != 0 || a.chan_list == nullptr) return cudaErrorInvalidValue;
    int per = (a.num_items + sm_count - 1) / sm_count;
    per = ((per + 3) / 4) * 4;
    if (per > FG) per = FG;
    if (per < 4) per = 4;
    const int grid = (a.num_items + per - 1) / per;
    cudaError_t e = cudaFuncSetAttribute(rx_ssb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    rx_ssb_tc_kernel<<<grid, NTHREADS, sizeof(Smem), stream>>>(a, per, dec_c, hil_ci, hil_cq);
    return cudaGetLastError();
}

bool rx_ssb_tc_available() { return true; }

}  // namespace uhsdr

#ifdef UHSDR_TC_PROF
// tools only (libuhsdr_b200_prof.so): read the role timers of the last launch, set the knock-out mask of the next ones
extern "C" int uhsdr_debug_tc_pause(int cycles) { return cudaMemcpyToSymbol(uhsdr::g_tc_pause, &cycles, sizeof(int)) == cudaSuccess ? 0 : -1; }
extern "C" int uhsdr_debug_tc_prof(unsigned long long *out, int knock)
{
    if (out && cudaMemcpyFromSymbol(out, uhsdr::g_tc_prof, sizeof(unsigned long long) * 32 * 4) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(uhsdr::g_tc_knock, &knock, sizeof(int)) != cudaSuccess) return -1;
    return 0;
}
#endif

#else   // UHSDR_EXACT: reference-order arithmetic has no tensor-core form

namespace uhsdr {
cudaError_t launch_rx_ssb_tc(const RxArgs &, int, int, int, int, cudaStream_t) { return cudaErrorNotSupported; }
bool rx_ssb_tc_available() { return false; }
}  // namespace uhsdr

#endif
