// Probe of tcgen05.mma kind::tf32 with K-major SWIZZLE_NONE operands and a skinny N = 8 (development): is the layout
// right, and what does a chain of 51 x 3 such MMAs (one Euler step of a 2N = 402 network) cost?
//   A [M][KT] (K contiguous)  -> smem: core matrix = 8 rows x 16 bytes; element (m, k) at
//                                 (k/4)*LBO + (m/8)*SBO + (m%8)*16 + (k%4)*4
//   B [N=8][KT]               -> (k/4)*128 + n*16 + (k%4)*4
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
constexpr int N = 8, KT = 408;
__device__ unsigned sa(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ unsigned long long desc(unsigned addr, unsigned lbo, unsigned sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3fffu);
    d |= (unsigned long long)((lbo >> 4) & 0x3fffu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;
    return d;                                            // swizzle mode 0 = none
}
template <int M>
__global__ void probe(const float *A, const float *B, float *D, long long *cycles, int swap, int reps, int terms) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_base;
    constexpr unsigned A_LBO = (M / 8) * 128, A_SBO = 128;
    float *As = reinterpret_cast<float *>(smem);
    float *Bs = reinterpret_cast<float *>(smem + (size_t)M * KT * 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * KT; i += blockDim.x) {
        const int m = i / KT, k = i % KT;
        As[((k / 4) * A_LBO + (m / 8) * A_SBO + (m % 8) * 16 + (k % 4) * 4) / 4] = A[i];
    }
    for (int i = tid; i < N * KT; i += blockDim.x) {
        const int n = i / KT, k = i % KT;
        Bs[((k / 4) * 128 + n * 16 + (k % 4) * 4) / 4] = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(sa(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base;
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    unsigned parity = 0;
    long long t0 = 0, t1 = 0;
    for (int r = 0; r < reps; ++r) {
        if (r == 1) t0 = clock64();
        if (tid == 0) {
            for (int t = 0; t < terms; ++t)
                for (int ks = 0; ks < KT / 8; ++ks) {
                    const unsigned a_addr = sa(As) + ks * 2 * A_LBO, b_addr = sa(Bs) + ks * 256;
                    const unsigned long long da = swap ? desc(a_addr, A_SBO, A_LBO) : desc(a_addr, A_LBO, A_SBO);
                    const unsigned long long db = swap ? desc(b_addr, 128, 128) : desc(b_addr, 128, 128);
                    const unsigned accum = (t | ks) ? 1u : 0u;
                    asm volatile(
                        "{ .reg .pred p; setp.ne.b32 p, %4, 0;"
                        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                        ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sa(&bar)) : "memory");
        }
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(sa(&bar)), "r"(parity) : "memory");
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // read the accumulator like the epilogue of a step would: 8 columns of this warp's 32 lanes
        unsigned v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tb + ((unsigned)((warp % 4) * 32) << 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (r == reps - 1)
            for (int c = 0; c < N; ++c) D[tid * N + c] = __uint_as_float(v[c]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) *cycles = reps > 1 ? (t1 - t0) / (reps - 1) : 0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}
template <int M>
void run(const char *name) {
    static float hA[M * KT], hB[N * KT], hD[128 * N];
    static double ref[M * N];
    srand(1);
    // tf32-exact inputs: small integers / 8
    for (int i = 0; i < M * KT; ++i) hA[i] = (float)(rand() % 33 - 16) / 8.f;
    for (int i = 0; i < N * KT; ++i) hB[i] = (float)(rand() % 17 - 8) / 4.f;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < KT; ++k) s += (double)hA[m * KT + k] * hB[n * KT + k]; ref[m * N + n] = s; }
    float *dA, *dB, *dD; long long *dc, hc;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD)); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    const int smem = M * KT * 4 + N * KT * 4;
    cudaFuncSetAttribute(probe<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // (LBO = distance of the core matrices along K, SBO = along M/N: the swapped reading faults)
    for (int swap = 0; swap < 1; ++swap)
        for (int terms = 1; terms <= 3; terms += 2) {
            cudaMemset(dD, 0xff, sizeof(hD));
            probe<M><<<1, 128, smem>>>(dA, dB, dD, dc, swap, 21, terms);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
            double err = 0; int bad = 0;
            // which TMEM lane holds row m (M = 64 uses a different lane map than M = 128)
            int lane_of[M];
            for (int m = 0; m < M; ++m) {
                lane_of[m] = -1;
                for (int l = 0; l < 128 && lane_of[m] < 0; ++l) {
                    bool ok = true;
                    for (int c = 0; c < N; ++c) ok = ok && fabs(hD[l * N + c] - terms * ref[m * N + c]) < 1e-3;
                    if (ok) lane_of[m] = l;
                }
                if (lane_of[m] < 0) ++bad; else { double d = 0; for (int c = 0; c < N; ++c) d = fmax(d, fabs(hD[lane_of[m] * N + c] - terms * ref[m * N + c])); if (d > err) err = d; }
            }
            if (terms == 1) { printf("%s: TMEM lane of rows 0, 1, 16, 17, 32, 63: %d %d %d %d %d %d\n", name, lane_of[0], lane_of[1], lane_of[16 % M], lane_of[17 % M], lane_of[32 % M], lane_of[63 % M]); }
            printf("%s swap=%d terms=%d: %s, max abs err %g, rows not found %d of %d, %lld cycles per step (%d MMAs + commit + wait + tcgen05.ld + sync)\n",
                   name, swap, terms, cudaGetErrorString(e), err, bad, M, hc, terms * KT / 8);
            if (e != cudaSuccess) exit(1);
        }
}
int main() {
    run<128>("M=128");
    run<64>("M=64");
    return 0;
}
