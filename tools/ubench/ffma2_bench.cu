// Microbenchmark: issue rate of fma.rn.f32x2 (FFMA2) vs FFMA in the K1 contraction pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { unsigned long long v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); return v; }
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long w, float x) {
    unsigned long long xx; asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(xx));
}
constexpr int NC = 14;
// MODE 0: FFMA2 (3 pairs + 1 single row, 4 stimuli), x from registers; MODE 1: plain FFMA 7 rows x 4; MODE 2: FFMA2 with x from LDS.128
template <int MODE>
__global__ void k(const float *w, const float *x, float *out, long long *cyc, int iters) {
    __shared__ float4 xs[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) xs[i] = make_float4(x[i & 63], x[(i + 1) & 63], x[(i + 2) & 63], x[(i + 3) & 63]);
    __syncthreads();
    unsigned long long wp[3][NC]; float ws[NC]; float wf[7][NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int q = 0; q < 3; ++q) wp[q][c] = pack2(w[(2 * q) * 448 + c * 32 + threadIdx.x], w[(2 * q + 1) * 448 + c * 32 + threadIdx.x]);
        ws[c] = w[6 * 448 + c * 32 + threadIdx.x];
#pragma unroll
        for (int t = 0; t < 7; ++t) wf[t][c] = w[t * 448 + c * 32 + threadIdx.x];
    }
    float xr[4] = {x[0], x[1], x[2], x[3]};
    unsigned long long ap[3][4]; float as[4]; float af[7][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) { for (int q = 0; q < 3; ++q) ap[q][b] = 0; as[b] = 0; for (int t = 0; t < 7; ++t) af[t][b] = 0; }
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float xv[4];
            if (MODE == 2) { float4 v = xs[(c * 32 + lane + it) & 511]; xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w; }
            else { xv[0] = xr[0]; xv[1] = xr[1]; xv[2] = xr[2]; xv[3] = xr[3]; }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (MODE == 1) {
#pragma unroll
                    for (int t = 0; t < 7; ++t) af[t][b] = fmaf(wf[t][c], xv[b], af[t][b]);
                } else {
#pragma unroll
                    for (int q = 0; q < 3; ++q) ffma2(ap[q][b], wp[q][c], xv[b]);
                    as[b] = fmaf(ws[c], xv[b], as[b]);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        for (int q = 0; q < 3; ++q) { float lo, hi; unpack2(ap[q][b], lo, hi); s += lo + hi; }
        s += as[b];
        for (int t = 0; t < 7; ++t) s += af[t][b];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float *w, *x, *out; long long *cyc;
    cudaMalloc(&w, 7 * 448 * 4 + 4096); cudaMalloc(&x, 4096); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    cudaMemset(w, 0, 7 * 448 * 4 + 4096); cudaMemset(x, 0, 4096);
    const int iters = 2000;
    for (int warps : {4, 8, 12, 16}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, 32 * warps>>>(w, x, out, cyc, iters);
                if (mode == 1) k<1><<<148, 32 * warps>>>(w, x, out, cyc, iters);
                if (mode == 2) k<2><<<148, 32 * warps>>>(w, x, out, cyc, iters);
                cudaDeviceSynchronize();
            }
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps/SM %2d mode %d (%s): %.1f cycles per half-panel pass (392 FMA-lanes-cycles per warp; floor %d)  err=%s\n", warps, mode,
                   mode == 0 ? "FFMA2 regs" : mode == 1 ? "FFMA regs" : "FFMA2 + LDS.128", (double)h / iters, 392 * warps / 4, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
