// Bring-up test for the tensor-core decimator of rx_ssb_tc2.cu:
//   y[m][col] = sum_k c[k] * x[col][4m - 82 + k]      (arm_fir_decimate_f32 convention, 83 taps, M = 4)
// as D[128 outputs][64 columns = 32 channels x {I, Q}] += A[128 x 16] * B[64 x 16]^T over 38 k-steps, with
//   A = sliding 128-row window of a strided Toeplitz table G[r][q] = c[q - 4 r + 578] stored with linear rows
//       (offset = (q / 8) * PLANE + r * 16 B + (q % 8) * 2 B; the window moves 4 rows = 64 B per k-step, so the
//       descriptor start address is NOT aligned to a 128-byte core matrix)
//   B = sample ring in the canonical K-major no-swizzle layout, I and Q channel groups contiguous (N = 64)
// and the accumulation split over two commits (incremental accumulation into a live TMEM accumulator).
// Prints the worst error against an fp64 reference, the lane that holds each row, and cycles per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_dec_toeplitz umma_dec_toeplitz.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int NT = 83, M = 128, N = 64, KSTEPS = 38, GROWS = 276, WIN = 608, R0 = 148, C0 = 578;
constexpr int PLANE = GROWS * 16;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo, unsigned sbo)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;
    return d;
}

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

struct Smem {
    alignas(128) unsigned char g[2][2 * PLANE];                 // c1 / c2 tables, linear rows
    alignas(128) unsigned char x[2][8 * (WIN / 8) * 128];       // x1 / x2: [group 0..7][time/8][col%8][time%8] bf16
    alignas(8) unsigned long long bar;
    unsigned tmem_base;
};

__global__ void __launch_bounds__(128, 1)
k(const float *c, const float *d, int LD, float *dump, volatile long long *cycles, int reps)
{
    extern __shared__ __align__(128) unsigned char raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < GROWS * 16; i += 128) {
        const int r = i >> 4, q = i & 15;
        const int kidx = q - 4 * r + C0;
        const float cv = (kidx >= 0 && kidx < NT) ? c[kidx] : 0.0f;
        const __nv_bfloat16 c1 = __float2bfloat16_rn(cv);
        const __nv_bfloat16 c2 = __float2bfloat16_rn(cv - __bfloat162float(c1));
        const int off = (q >> 3) * PLANE + r * 16 + (q & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(sm.g[0] + off) = c1;
        *reinterpret_cast<__nv_bfloat16 *>(sm.g[1] + off) = c2;
    }
    // window slot s holds sample n = s - 96 of column col; d is [64][LD] with sample n at index n + 96
    for (int i = tid; i < 64 * WIN; i += 128) {
        const int col = i / WIN, s = i % WIN;
        const float xv = d[col * LD + s];
        const __nv_bfloat16 x1 = __float2bfloat16_rn(xv);
        const __nv_bfloat16 x2 = __float2bfloat16_rn(xv - __bfloat162float(x1));
        const int off = (col >> 3) * (WIN / 8) * 128 + (s >> 3) * 128 + (col & 7) * 16 + (s & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(sm.x[0] + off) = x1;
        *reinterpret_cast<__nv_bfloat16 *>(sm.x[1] + off) = x2;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    const unsigned idesc32 = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(32 >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    const unsigned g0 = smem_u32(sm.g[0]), g1 = smem_u32(sm.g[1]), x0 = smem_u32(sm.x[0]), x1a = smem_u32(sm.x[1]);
    const unsigned sbo_b = (WIN / 8) * 128;
    long long t0 = 0, t1 = 0;
    unsigned phase = 0;
    auto wait = [&]() {
        asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}"
                     ::"r"(smem_u32(&sm.bar)), "r"(phase) : "memory");
        phase ^= 1;
    };
    // ---- correctness: two halves of the k range, committed separately ----
    for (int half = 0; half < 2; half++) {
        if (tid == 0) {
            for (int kk = half * 19; kk < half * 19 + 19; kk++) {
                const unsigned arow = (unsigned)(R0 - 4 * kk) * 16;
                const unsigned bt = (unsigned)(2 * kk) * 128;
                const unsigned long long a1 = make_desc(g0 + arow, PLANE, 128), a2 = make_desc(g1 + arow, PLANE, 128);
                const unsigned long long b1 = make_desc(x0 + bt, 128, sbo_b), b2 = make_desc(x1a + bt, 128, sbo_b);
                umma_bf16(tmem, a1, b1, idesc, kk > 0 ? 1u : 0u);
                umma_bf16(tmem, a2, b1, idesc, 1u);
                umma_bf16(tmem, a1, b2, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.bar)) : "memory");
        }
        wait();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
    }
    {
        unsigned v[64];
        const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                     "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                       "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                       "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
                       "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]),
                       "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]),
                       "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 64; j++) dump[tid * 64 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- timing.  kind 0: 24 MMAs M=128 N=64; 1: 48 of them; 2: 24 MMAs M=128 N=32; 3: 48; 4: 24 MMAs M=64 N=32; 5: 48.
    // Descriptors advance by additions in a fully unrolled loop (the issue loop of the real kernel).
    const unsigned idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(32 >> 3) << 17) | ((unsigned)(64 >> 4) << 24);
    for (int rep = 0; rep < reps; rep++) {
        const int kind = rep % 6;
        if (warp == 0 && elect_one()) {
            const unsigned id = (kind < 2) ? idesc : (kind < 4) ? idesc32 : idesc64;
            const unsigned long long a1_0 = make_desc(g0 + R0 * 16, PLANE, 128), a2_0 = make_desc(g1 + R0 * 16, PLANE, 128);
            const unsigned long long b1_0 = make_desc(x0, 128, sbo_b), b2_0 = make_desc(x1a, 128, sbo_b);
            t0 = clock64(); cycles[2 * reps] = t0;
            if (kind & 1) {
#pragma unroll
                for (int kk = 0; kk < 16; kk++) {
                    umma_bf16(tmem, a1_0 - 4 * kk, b1_0 + 16 * kk, id, 1u);
                    umma_bf16(tmem, a2_0 - 4 * kk, b1_0 + 16 * kk, id, 1u);
                    umma_bf16(tmem, a1_0 - 4 * kk, b2_0 + 16 * kk, id, 1u);
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 8; kk++) {
                    umma_bf16(tmem, a1_0 - 4 * kk, b1_0 + 16 * kk, id, 1u);
                    umma_bf16(tmem, a2_0 - 4 * kk, b1_0 + 16 * kk, id, 1u);
                    umma_bf16(tmem, a1_0 - 4 * kk, b2_0 + 16 * kk, id, 1u);
                }
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.bar)) : "memory");
            cycles[reps + rep] = t1 - t0;       // issue time alone
        }
        wait();
        if (tid == 0) { t1 = clock64(); cycles[rep] = t1 - cycles[2 * reps]; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main()
{
    const int LD = WIN;
    std::vector<float> c(NT), d(64 * LD);
    srand(7);
    for (auto &v : c) v = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
    for (auto &v : d) v = (rand() / (float)RAND_MAX - 0.5f) * 20000.0f;
    float *dc, *dd, *ddump; long long *dcyc;
    const int reps = 36;
    cudaMalloc(&dc, c.size() * 4); cudaMalloc(&dd, d.size() * 4); cudaMalloc(&ddump, 128 * 64 * 4); cudaMalloc(&dcyc, (2 * reps + 1) * 8);
    cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dd, d.data(), d.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(ddump, 0, 128 * 64 * 4); cudaMemset(dcyc, 0, (2 * reps + 1) * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    k<<<1, 128, sizeof(Smem)>>>(dc, dd, LD, ddump, dcyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s, smem %zu B\n", cudaGetErrorString(e), sizeof(Smem));
    if (e != cudaSuccess) return 1;
    std::vector<float> dump(128 * 64); std::vector<long long> cyc(2 * reps + 1);
    cudaMemcpy(dump.data(), ddump, dump.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc.data(), dcyc, (2 * reps + 1) * 8, cudaMemcpyDeviceToHost);
    double worst = 0, rms = 0, refrms = 0, scale = 0;
    int bad_rows = 0;
    for (int m = 0; m < M; m++) {
        double rowworst = 0;
        for (int col = 0; col < N; col++) {
            double s = 0;
            for (int kx = 0; kx < NT; kx++) s += (double)c[kx] * (double)d[col * LD + (4 * m - 82 + kx) + 96];
            const double er = dump[m * 64 + col] - s;
            rowworst = fmax(rowworst, fabs(er)); rms += er * er; refrms += s * s; scale = fmax(scale, fabs(s));
        }
        worst = fmax(worst, rowworst);
        if (rowworst > 1.0) { if (bad_rows < 8) printf("row %d: worst err %.3e\n", m, rowworst); bad_rows++; }
    }
    printf("rows with error > 1: %d of %d\n", bad_rows, M);
    printf("worst |err| %.3e (full scale %.3e, rel %.3e), SNR %.1f dB\n", worst, scale, worst / scale, 10 * log10(refrms / rms));
    const char *names[6] = {"24 x M128 N64", "48 x M128 N64", "24 x M128 N32", "48 x M128 N32", "24 x M64 N32", "48 x M64 N32"};
    double tot[6] = {0}, iss[6] = {0}; int cnt[6] = {0};
    for (int r = 6; r < reps; r++) { tot[r % 6] += cyc[r]; iss[r % 6] += cyc[reps + r]; cnt[r % 6]++; }
    for (int kd = 0; kd < 6; kd++) printf("%-14s total %.0f cycles (issue alone %.0f)\n", names[kd], tot[kd] / cnt[kd], iss[kd] / cnt[kd]);
    for (int kd = 0; kd < 6; kd += 2) printf("marginal cycles per MMA (%s): %.1f\n", names[kd] + 5, (tot[kd + 1] / cnt[kd + 1] - tot[kd] / cnt[kd]) / 24.0);
    return 0;
}
