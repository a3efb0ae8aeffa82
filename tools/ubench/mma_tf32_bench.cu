// Microbenchmark: issue rate of the legacy mma.sync.m16n8k8 TF32 path on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32_bench mma_tf32_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NACC>
__global__ void k(float *out, int iters, long long *cyc) {
    float c[NACC][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 3 + i;
    for (int i = 0; i < 2; ++i) b[i] = threadIdx.x * 5 + i;
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int j = 0; j < NACC; ++j) mma_tf32(c[j], a, b);
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    for (int warps : {4, 8, 16, 32}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<16><<<148, 32 * warps>>>(out, iters, cyc); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<16><<<148, 32 * warps>>>(out, iters, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double flop = 148.0 * warps * iters * 16 * 2048.0;
        printf("warps/SM %2d: %.3f ms, %.1f TFLOP/s tf32 (m16n8k8), %.2f cycles per mma per warp-scheduler slot, err=%s\n", warps, ms, flop / ms * 1e-9,
               (double)h / (iters * 16.0) / (warps / 4.0 > 1 ? 1 : 1), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
