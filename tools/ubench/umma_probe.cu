// Probe of tcgen05.mma kind::tf32 with MN-major SWIZZLE_128B operands (development): one CTA, one K = 8 step.
// A is [K=8][M=128] (M contiguous), B is [K=8][N=32] (N contiguous), both written to shared memory in the layout a
// TMA box {32 floats, 8 rows} with the 128-byte / 32-byte-atom swizzle produces (the only swizzled MN-major layout
// 32-bit operands may use: SWIZZLE_128B_BASE32B); D = A^T B is read back from TMEM and compared.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int M = 128, N = 32, K = 8;
__device__ unsigned sa(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ unsigned long long desc(unsigned addr, unsigned lbo, unsigned sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3fffu);
    d |= (unsigned long long)((lbo >> 4) & 0x3fffu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;
    d |= 1ull << 61;                                    // SWIZZLE_128B_BASE32B
    return d;
}
__global__ void probe(const float *A, const float *B, float *D, unsigned idesc, int variant) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = raw + ((1024u - (sa(raw) & 1023u)) & 1023u);
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_base;
    float *As = reinterpret_cast<float *>(smem);              // 4 boxes of [8 rows][32 floats]
    float *Bs = reinterpret_cast<float *>(smem + 4096);       // 1 box
    const int tid = threadIdx.x, warp = tid >> 5;
    // box b, row k, element e (0..31): byte offset b*1024 + k*128 + ((e/8) ^ (k & 3))*32 + (e%8)*4
    for (int i = tid; i < K * M; i += blockDim.x) {
        const int k = i / M, m = i % M, b = m / 32, e = m % 32;
        As[(b * 1024 + k * 128 + (((e / 8) ^ (k & 3)) * 32) + (e % 8) * 4) / 4] = A[k * M + m];
    }
    for (int i = tid; i < K * N; i += blockDim.x) {
        const int k = i / N, e = i % N;
        Bs[(k * 128 + (((e / 8) ^ (k & 3)) * 32) + (e % 8) * 4) / 4] = B[k * N + e];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(sa(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base;
    if (variant == 1) {          // TMEM store / load round trip only
        unsigned v = tid * 1000;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb + ((unsigned)(warp * 32) << 16)), "r"(v) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (tid == 0) {
        asm volatile(
            "{ .reg .pred p; setp.ne.b32 p, %4, 0;"
            " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
            ::"r"(tb), "l"(desc(sa(As), 1024, 512)), "l"(desc(sa(Bs), 1024, 512)), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sa(&bar)) : "memory");
    }
    if (variant != 1) {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(sa(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; ++c) {
        unsigned r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tb + ((unsigned)(warp * 32) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        D[tid * N + c] = variant == 1 ? (float)r : __uint_as_float(r);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb) : "memory");
}
int main(int argc, char **argv) {
    float hA[K * M], hB[K * N], hD[M * N], ref[M * N];
    for (int i = 0; i < K * M; ++i) hA[i] = (float)((i * 7) % 13 - 6);
    for (int i = 0; i < K * N; ++i) hB[i] = (float)((i * 5) % 11 - 5);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += hA[k * M + m] * hB[k * N + n]; ref[m * N + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    const unsigned base = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    const unsigned variants[3] = {base | (1u << 15) | (1u << 16), 0, base | (1u << 15) | (1u << 16)};
    for (int v = 0; v < 2; ++v) {
        cudaMemset(dD, 0xff, sizeof(hD));
        probe<<<1, 128, 8192 + 1024>>>(dA, dB, dD, variants[v], v);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        if (v == 1) { printf("variant 1 (tmem st/ld): err %s, D[0]=%g D[33*N]=%g D[127*N]=%g (expect 0, 33000, 127000)\n", cudaGetErrorString(e), hD[0], hD[33 * N], hD[127 * N]); continue; }
        double err = 0; int bad = 0;
        for (int i = 0; i < M * N; ++i) { double d = fabs(hD[i] - ref[i]); if (d > err) err = d; if (d > 1e-3) ++bad; }
        printf("variant %d (mma MN-major SW128/32B atom): err %s, max abs err %g, bad %d of %d; D[0..3] = %g %g %g %g ref %g %g %g %g\n", v,
               cudaGetErrorString(e), err, bad, M * N, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
