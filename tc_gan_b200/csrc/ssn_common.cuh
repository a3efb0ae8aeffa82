// Shared definitions for the SSN kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/ssnode.h"

namespace ssn {

// ---- transfer function, one source for host (double ABI helpers) and device ----
// Reference: tc_gan/ext/ssnode.c:21-53 and tc_gan/ssnode.py:125-149.

template <typename T> __host__ __device__ inline T pow_t(T a, T b);
template <> __host__ __device__ inline double pow_t<double>(double a, double b) { return pow(a, b); }
template <> __host__ __device__ inline float pow_t<float>(float a, float b) { return powf(a, b); }
template <typename T> __host__ __device__ inline T tanh_t(T a);
template <> __host__ __device__ inline double tanh_t<double>(double a) { return tanh(a); }
template <> __host__ __device__ inline float tanh_t<float>(float a) { return tanhf(a); }
template <typename T> __host__ __device__ inline T cosh_t(T a);
template <> __host__ __device__ inline double cosh_t<double>(double a) { return cosh(a); }
template <> __host__ __device__ inline float cosh_t<float>(float a) { return coshf(a); }

// Constants of f for one (k, n, r_soft, r_hard): precomputed on the host in
// double and rounded once, so device float code does no divisions per call.
template <typename T>
struct IoConst {
    int io_type;
    T k, n, v0, r_soft, span;      // span = r_hard - r_soft
    T lin_slope;                   // k n v0^(n-1)
    T tanh_scale;                  // n r_soft / (span v0)
    T gain_hi;                     // n r_soft / v0
    T nk;                          // n k
    int n_int;                     // floor(n), capped at 8: v^n = v^n_int * v^n_frac (io_power_fast)
    T n_frac;
};

template <typename T>
inline IoConst<T> make_io_const(int io_type, double k, double n, double r_soft, double r_hard) {
    IoConst<T> c;
    const double v0 = pow(r_soft / k, 1.0 / n);
    c.io_type = io_type;
    c.k = (T)k; c.n = (T)n; c.v0 = (T)v0; c.r_soft = (T)r_soft;
    c.span = (T)(r_hard - r_soft);
    c.lin_slope = (T)(k * pow(v0, n - 1.0) * n);
    c.tanh_scale = (T)(n * r_soft / ((r_hard - r_soft) * v0));
    c.gain_hi = (T)(n * r_soft / v0);
    c.nk = (T)(n * k);
    c.n_int = n >= 1.0 ? (n < 8.0 ? (int)n : 8) : 0;
    c.n_frac = (T)(n - c.n_int);
    return c;
}

// f(v)
template <typename T>
__host__ __device__ inline T io_eval(const IoConst<T> &c, T v) {
    if (v <= (T)0) return (T)0;
    if (c.io_type == SSN_IO_POWER || v <= c.v0) return c.k * pow_t<T>(v, c.n);
    if (c.io_type == SSN_IO_LINEAR) return c.r_soft + c.lin_slope * (v - c.v0);
    return c.r_soft + c.span * tanh_t<T>(c.tanh_scale * (v - c.v0));
}

// f'(v) as tc_gan/gradient_expressions/SS_grad.py:78-99 defines it.
template <typename T>
__host__ __device__ inline T io_gain(const IoConst<T> &c, T v) {
    const T vc = v > (T)0 ? v : (T)0;
    if (c.io_type == SSN_IO_POWER) return c.nk * pow_t<T>(vc, c.n - (T)1);
    if (c.io_type == SSN_IO_LINEAR) {
        const T vv = vc < c.v0 ? vc : c.v0;
        return c.nk * pow_t<T>(vv, c.n - (T)1);
    }
    if (vc <= c.v0) return c.nk * pow_t<T>(vc, c.n - (T)1);
    const T ch = cosh_t<T>(c.tanh_scale * (vc - c.v0));
    return c.gain_hi / (ch * ch);
}

// ---- W(z; J, D, S) -------------------------------------------------------------
// W[aN+i, bN+j] = s_b exp(-(x_i-x_j)^2 / (2 S_ab^2)) (J_ab + D_ab z), x = linspace(-.5,.5,N)
// (tc_gan/weight_gen.py:6-26).  inv2s2[ab] = 1/(2 S_ab^2), sJ/sD carry the sign s_b.
struct WeightConst {
    float sJ[4], sD[4], inv2s2[4], invS3[4];
    float dx;            // 1/(N-1)
};

inline WeightConst make_weight_const(const ssn_jds &p, int n_sites) {
    WeightConst w;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int ab = a * 2 + b;
            const double s = b == 0 ? 1.0 : -1.0;
            w.sJ[ab] = (float)(s * p.J[ab]);
            w.sD[ab] = (float)(s * p.D[ab]);
            w.inv2s2[ab] = (float)(1.0 / (2.0 * p.S[ab] * p.S[ab]));
            w.invS3[ab] = (float)(1.0 / (p.S[ab] * p.S[ab] * p.S[ab]));
        }
    w.dx = n_sites > 1 ? (float)(1.0 / (n_sites - 1)) : 0.f;
    return w;
}

// Gaussian profile table gtab[ab * n_sites + d] = exp(-(d dx)^2 / (2 S_ab^2)), d = |i - j|
// site distance: 4 N accurate expf per CTA instead of one per W element.
__device__ __forceinline__ void build_profile_table(const WeightConst &wc, int n_sites, float *gtab,
                                                    int tid, int nthreads) {
    for (int idx = tid; idx < 4 * n_sites; idx += nthreads) {
        const int ab = idx / n_sites, d = idx - ab * n_sites;
        const float x = (float)d * wc.dx;
        gtab[idx] = expf(-x * x * wc.inv2s2[ab]);
    }
}

// W element (i, j) of a network from its z element; also returns block index and profile.
__device__ __forceinline__ float weight_from_z(const WeightConst &wc, const float *gtab, int n_sites,
                                               int i, int j, float z) {
    const int a = i >= n_sites, b = j >= n_sites;
    const int ab = a * 2 + b;
    int d = (i - a * n_sites) - (j - b * n_sites);
    d = d < 0 ? -d : d;
    return gtab[ab * n_sites + d] * fmaf(wc.sD[ab], z, wc.sJ[ab]);
}

// ---- host-side bookkeeping -------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_cuda(cudaError_t e, const char *what);       // 0 or 1000 + e (and records text)

// Scope around one kernel launch: with ssn_profile_enable(1) it records CUDA events on the launching
// stream before and after, for ssn_profile_read (per-kernel device time inside bench.py); otherwise a no-op.
class KernelTimer {
public:
    KernelTimer(const char *name, cudaStream_t stream);
    ~KernelTimer();
    KernelTimer(const KernelTimer &) = delete;
    KernelTimer &operator=(const KernelTimer &) = delete;
private:
    const char *name_;
    cudaStream_t stream_;
    cudaEvent_t e0_ = nullptr, e1_ = nullptr;
};

#define SSN_CUDA(call)                                        \
    do {                                                      \
        int _rc = ::ssn::check_cuda((call), #call);           \
        if (_rc) return _rc;                                  \
    } while (0)

}  // namespace ssn
