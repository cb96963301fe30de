// Microbenchmark / bring-up test for the tensor-core Hilbert FIR of rx_ssb_tc.cu:
//   y[n][ch] = sum_k c[k] * d[ch][n - 198 + k]        (arm_fir_f32 convention, 199 taps)
// as D[M=64 outputs][N=32 channels] += A[64 x 16] * B[32 x 16]^T over 17 k-steps with
//   A = sliding 64-row window of a Toeplitz table G[r][q] = c[q - r + 254] (constant, shared memory)
//   B = channel-major sample ring in the canonical K-major no-swizzle layout
// bf16 split arithmetic: x = x1 + x2, c = c1 + c2, D = c1*x1 + c2*x1 + c1*x2  (fp32 accumulate).
// Prints the TMEM lane that holds each output row, the worst error against an fp64 reference and
// the cycles per tcgen05.mma when 102 of them (one chunk of the real kernel) are issued back to back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_toeplitz umma_toeplitz.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int NT = 199, M = 64, N = 32, KSTEPS = 17, GROWS = 328, RING = 336;   // ring slots (multiple of 16)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo, unsigned sbo)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;      // descriptor version (Blackwell)
    return d;             // layout_type 0 = no swizzle
}

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct Smem {
    alignas(128) __nv_bfloat16 g[2][GROWS * 16];       // c1 / c2 Toeplitz tables: [row group][khalf][row%8][8]
    alignas(128) __nv_bfloat16 x[2][4 * (RING / 8) * 64];   // x1 / x2 rings: [chgroup][time/8][ch%8][time%8]
    alignas(8) unsigned long long bar;
    unsigned tmem_base;
};

__global__ void __launch_bounds__(128, 1)
k(const float *c, const float *d, int T, float *dump, long long *cycles, int reps)
{
    extern __shared__ __align__(128) unsigned char raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- Toeplitz tables ----
    for (int i = tid; i < GROWS * 16; i += 128) {
        const int r = i >> 4, q = i & 15;
        const int kidx = q - r + 254;
        const float cv = (kidx >= 0 && kidx < NT) ? c[kidx] : 0.0f;
        const __nv_bfloat16 c1 = __float2bfloat16_rn(cv);
        const __nv_bfloat16 c2 = __float2bfloat16_rn(cv - __bfloat162float(c1));
        const int off = (r >> 3) * 128 + (q >> 3) * 64 + (r & 7) * 8 + (q & 7);
        sm.g[0][off] = c1; sm.g[1][off] = c2;
    }
    // ---- data ring: time slot s holds sample p = s - 208 (window of chunk 0 = slots [0, 272)) ----
    for (int i = tid; i < 32 * RING; i += 128) {
        const int ch = i / RING, s = i % RING;
        const int p = s - 208;
        const float xv = (p >= -198 && p < T) ? d[ch * (T + 198) + p + 198] : 0.0f;
        const __nv_bfloat16 x1 = __float2bfloat16_rn(xv);
        const __nv_bfloat16 x2 = __float2bfloat16_rn(xv - __bfloat162float(x1));
        const int off = (ch >> 3) * (RING / 8) * 64 + (s >> 3) * 64 + (ch & 7) * 8 + (s & 7);
        sm.x[0][off] = x1; sm.x[1][off] = x2;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    const unsigned g0 = smem_u32(sm.g[0]), g1 = smem_u32(sm.g[1]), x0 = smem_u32(sm.x[0]), x1a = smem_u32(sm.x[1]);
    const unsigned sbo_b = (RING / 8) * 128;
    long long t0 = 0, t1 = 0;
    unsigned phase = 0;
    for (int rep = 0; rep < reps; rep++) {
        if (tid == 0) {
            t0 = clock64();
            for (int pass = 0; pass < (rep == 0 ? 1 : 2); pass++) {      // timing reps: 2 x 51 = 102 MMAs (I and Q of the real kernel)
                for (int kk = 0; kk < KSTEPS; kk++) {
                    const unsigned arow = (unsigned)(264 - 16 * kk) * 32;       // byte offset of G row R(kk)
                    const unsigned bt = (unsigned)(16 * kk / 8) * 128;          // byte offset of time slot 16kk
                    const unsigned long long a1 = make_desc(g0 + arow, 128, 256), a2 = make_desc(g1 + arow, 128, 256);
                    const unsigned long long b1 = make_desc(x0 + bt, 128, sbo_b), b2 = make_desc(x1a + bt, 128, sbo_b);
                    umma_bf16(tmem, a1, b1, idesc, (kk > 0 || pass > 0) ? 1u : 0u);
                    umma_bf16(tmem, a2, b1, idesc, 1u);
                    umma_bf16(tmem, a1, b2, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.bar)) : "memory");
        }
        // everybody waits for the MMAs of this repetition
        asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}"
                     ::"r"(smem_u32(&sm.bar)), "r"(phase) : "memory");
        phase ^= 1;
        if (tid == 0) { t1 = clock64(); if (rep > 0) cycles[rep] = t1 - t0; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (rep == 0) {
            // dump all 128 lanes x 32 columns of the accumulator
            unsigned v[32];
            const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; j++) dump[tid * 32 + j] = __uint_as_float(v[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main()
{
    const int T = 64;
    std::vector<float> c(NT), d(32 * (T + 198));
    srand(7);
    for (auto &v : c) v = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
    for (auto &v : d) v = (rand() / (float)RAND_MAX - 0.5f) * 20000.0f;
    float *dc, *dd, *ddump; long long *dcyc;
    const int reps = 21;
    cudaMalloc(&dc, c.size() * 4); cudaMalloc(&dd, d.size() * 4); cudaMalloc(&ddump, 128 * 32 * 4); cudaMalloc(&dcyc, reps * 8);
    cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dd, d.data(), d.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(ddump, 0, 128 * 32 * 4); cudaMemset(dcyc, 0, reps * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    k<<<1, 128, sizeof(Smem)>>>(dc, dd, T, ddump, dcyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s, smem %zu B\n", cudaGetErrorString(e), sizeof(Smem));
    if (e != cudaSuccess) return 1;
    std::vector<float> dump(128 * 32); std::vector<long long> cyc(reps);
    cudaMemcpy(dump.data(), ddump, dump.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc.data(), dcyc, reps * 8, cudaMemcpyDeviceToHost);
    // reference
    std::vector<double> ref(M * N);
    double scale = 0;
    for (int n = 0; n < M; n++)
        for (int ch = 0; ch < N; ch++) {
            double s = 0;
            for (int kx = 0; kx < NT; kx++) s += (double)c[kx] * (double)d[ch * (T + 198) + (n - 198 + kx) + 198];
            ref[n * N + ch] = s; scale = fmax(scale, fabs(s));
        }
    // which TMEM lane holds output row n?
    int lane_of[M]; double worst = 0, rms = 0, refrms = 0;
    for (int n = 0; n < M; n++) {
        int best = -1; double beste = 1e30;
        for (int l = 0; l < 128; l++) {
            double er = 0;
            for (int ch = 0; ch < N; ch++) er = fmax(er, fabs(dump[l * 32 + ch] - ref[n * N + ch]));
            if (er < beste) { beste = er; best = l; }
        }
        lane_of[n] = best; worst = fmax(worst, beste);
        for (int ch = 0; ch < N; ch++) { double er = dump[best * 32 + ch] - ref[n * N + ch]; rms += er * er; refrms += ref[n * N + ch] * ref[n * N + ch]; }
    }
    printf("lane of rows 0..63:");
    for (int n = 0; n < M; n++) printf(" %d", lane_of[n]);
    printf("\nworst |err| %.3e (full scale %.3e, rel %.3e), SNR %.1f dB\n", worst, scale, worst / scale, 10 * log10(refrms / rms));
    double avg = 0; for (int r = 1; r < reps; r++) avg += cyc[r];
    avg /= (reps - 1);
    printf("102 MMAs (M=64,N=32,K=16 bf16, smem x smem): %.0f cycles incl. commit+wait = %.1f cycles per MMA\n", avg, avg / 102.0);
    return 0;
}
