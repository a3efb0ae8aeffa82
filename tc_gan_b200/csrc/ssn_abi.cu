// extern "C" surface of libssnode.so (see include/ssnode.h): the reference ABI
// symbols of tc_gan/ext/ssnode.c plus the batched entry points, host staging and
// per-thread CUDA state.  No CPU fallback anywhere: without a usable GPU every
// solver call returns 1000 + cudaError_t and ssn_last_error() says why.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>
#include "ssn_launch.h"

namespace ssn {

// ---- error text, launch counter ------------------------------------------------
static thread_local char tl_error[512] = "";
static std::atomic<int> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return 1000 + (int)e;
}

// ---- work counters for the persistent kernels (one int per launch) ----------------
// A launch leases a slot of a per-device ring.  Device-pointer entry points are asynchronous, so a slot may
// still be in use by a kernel of another stream when the ring wraps: every slot carries an event recorded
// after the launch that used it, and the next lessee's stream waits on that event before its memset.
constexpr int COUNTER_RING = 256;
struct CounterSlot { cudaEvent_t done = nullptr; bool used = false, leased = false; };
struct CounterRing { int *base = nullptr; int pos = 0; CounterSlot slot[COUNTER_RING]; };
static std::mutex g_counter_mutex;
static CounterRing g_rings[64];

struct CounterLease {
    int *ptr = nullptr;
    int dev = -1, idx = -1;
    cudaStream_t stream = nullptr;
    int acquire(cudaStream_t st) {
        release();
        SSN_CUDA(cudaGetDevice(&dev));
        stream = st;
        std::lock_guard<std::mutex> lock(g_counter_mutex);
        CounterRing &ring = g_rings[dev & 63];
        if (!ring.base) SSN_CUDA(cudaMalloc(&ring.base, COUNTER_RING * sizeof(int)));
        for (int tries = 0; tries < COUNTER_RING; ++tries) {
            const int i = ring.pos++ % COUNTER_RING;
            CounterSlot &s = ring.slot[i];
            if (s.leased) continue;                       // another thread is between acquire and launch
            if (!s.done) SSN_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            if (s.used) SSN_CUDA(cudaStreamWaitEvent(st, s.done, 0));
            s.leased = true;
            idx = i;
            ptr = ring.base + i;
            return 0;
        }
        set_error("work-counter ring exhausted (%d launches being enqueued at once)", COUNTER_RING);
        return -1;
    }
    void release() {
        if (idx < 0) return;
        std::lock_guard<std::mutex> lock(g_counter_mutex);
        CounterSlot &s = g_rings[dev & 63].slot[idx];
        s.used = cudaEventRecord(s.done, stream) == cudaSuccess;
        s.leased = false;
        idx = -1;
    }
    ~CounterLease() { release(); }
};

// ---- optional per-kernel timing (ssn_profile_enable) -------------------------------
// CUDA events around each launch, on the stream the kernel is launched on; read back (and reset) with
// ssn_profile_read.  Off by default: two event records per launch are not free.
static std::atomic<int> g_profile{0};
struct TimedLaunch { const char *name; cudaEvent_t e0, e1; };
static std::mutex g_profile_mutex;
static std::vector<TimedLaunch> g_timed;

KernelTimer::KernelTimer(const char *name, cudaStream_t st) : name_(name), stream_(st) {
    if (!g_profile.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) { e0_ = e1_ = nullptr; return; }
    cudaEventRecord(e0_, stream_);
}
KernelTimer::~KernelTimer() {
    if (!e0_) return;
    cudaEventRecord(e1_, stream_);
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    g_timed.push_back({name_, e0_, e1_});
}

// ---- per-thread host staging ------------------------------------------------------
struct HostCtx {
    int device = -1;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    std::vector<void *> slot;
    std::vector<size_t> cap;
    int ensure_device() {
        int dev = 0;
        SSN_CUDA(cudaGetDevice(&dev));
        if (dev != device) {
            release();
            device = dev;
            SSN_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        }
        return 0;
    }
    int second_stream(cudaStream_t *out) {
        if (!stream2) SSN_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
        *out = stream2;
        return 0;
    }
    int get(int i, size_t bytes, void **out) {
        if ((int)slot.size() <= i) { slot.resize(i + 1, nullptr); cap.resize(i + 1, 0); }
        if (cap[i] < bytes) {
            if (slot[i]) cudaFree(slot[i]);
            slot[i] = nullptr; cap[i] = 0;
            SSN_CUDA(cudaMalloc(&slot[i], bytes));
            cap[i] = bytes;
        }
        *out = slot[i];
        return 0;
    }
    // pinned host staging (grown on demand): copies from it are truly asynchronous and do not serialise the
    // calling threads inside the driver the way copies from pageable memory do
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    int get_pinned(size_t bytes, void **out) {
        if (pinned_cap < bytes) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr; pinned_cap = 0;
            SSN_CUDA(cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault));
            pinned_cap = bytes;
        }
        *out = pinned;
        return 0;
    }
    void release() {
        for (void *p : slot) if (p) cudaFree(p);
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr; pinned_cap = 0;
        slot.clear(); cap.clear();
        if (stream) cudaStreamDestroy(stream);
        if (stream2) cudaStreamDestroy(stream2);
        stream = nullptr; stream2 = nullptr;
    }
    ~HostCtx() { release(); }
};
// Contexts outlive the threads that use them: the reference creates a new thread pool per find_fixed_points
// call (tc_gan/ssnode.py:455-460), and a fresh stream + device scratch + pinned buffer per new thread would
// cost more than the solves.  A thread borrows a context on first use and returns it when it exits.
static std::mutex g_ctx_mutex;
static std::vector<HostCtx *> g_ctx_free;
struct HostCtxHandle {
    HostCtx *p = nullptr;
    HostCtx &get() {
        if (!p) {
            std::lock_guard<std::mutex> lock(g_ctx_mutex);
            if (!g_ctx_free.empty()) { p = g_ctx_free.back(); g_ctx_free.pop_back(); }
            else p = new HostCtx();
        }
        return *p;
    }
    ~HostCtxHandle() {
        if (p) {
            std::lock_guard<std::mutex> lock(g_ctx_mutex);
            g_ctx_free.push_back(p);                 // never destroyed: no CUDA calls at thread or process exit
        }
    }
};
static thread_local HostCtxHandle tl_handle;
#define tl_ctx (tl_handle.get())

#define GET(i, T, n, var) T *var = nullptr; { void *_p; int _rc = tl_ctx.get(i, (size_t)(n) * sizeof(T), &_p); if (_rc) return _rc; var = (T *)_p; }

// ---- small conversion / weight kernels --------------------------------------------
__global__ void convert_f64_f32_kernel(const double *s, float *d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = (float)s[i];
}
__global__ void convert_f32_f64_kernel(const float *s, double *d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = (double)s[i];
}
static int grid_for(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 148 * 16); }
int launch_convert_f64_to_f32(const double *s, float *d, size_t n, cudaStream_t st) {
    if (!n) return 0;
    convert_f64_f32_kernel<<<grid_for(n), 256, 0, st>>>(s, d, n);
    count_launch();
    return check_cuda(cudaGetLastError(), "convert_f64_f32");
}
int launch_convert_f32_to_f64(const float *s, double *d, size_t n, cudaStream_t st) {
    if (!n) return 0;
    convert_f32_f64_kernel<<<grid_for(n), 256, 0, st>>>(s, d, n);
    count_launch();
    return check_cuda(cudaGetLastError(), "convert_f32_f64");
}

__global__ void generate_weight_kernel(int n_sites, const float *z, WeightConst wc, float *W, size_t total) {
    const int dim = 2 * n_sites;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(e % dim), i = (int)((e / dim) % dim);
        const int a = i >= n_sites, b = j >= n_sites, ab = a * 2 + b;
        const float d = (float)((i - a * n_sites) - (j - b * n_sites)) * wc.dx;
        W[e] = expf(-d * d * wc.inv2s2[ab]) * fmaf(wc.sD[ab], z[e], wc.sJ[ab]);
    }
}
int launch_generate_weight(int nz, int n_sites, const float *z, const ssn_jds &jds, float *W, cudaStream_t st) {
    const size_t total = (size_t)nz * 4 * n_sites * n_sites;
    if (!total) return 0;
    generate_weight_kernel<<<grid_for(total), 256, 0, st>>>(n_sites, z, make_weight_const(jds, n_sites), W, total);
    count_launch();
    return check_cuda(cudaGetLastError(), "generate_weight");
}

// FP32 FFMA peak probe: 8 independent FMA chains per thread, 1024 threads per SM-sized block.
__global__ void __launch_bounds__(1024) ffma_peak_kernel(float *out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;          // never true; keeps the chains alive
}

static int validate(const ssn_solver *sv, int nz, int nb, int n_sites) {
    if (!sv) { set_error("solver is NULL"); return -1; }
    if (sv->io_type < 0 || sv->io_type > 2) { set_error("unknown io_type %d", sv->io_type); return -1; }
    if (nz < 0 || nb < 0 || n_sites < 1) { set_error("bad shape nz=%d nb=%d n_sites=%d", nz, nb, n_sites); return -1; }
    return 0;
}

// ---- reference ABI: one network x one stimulus, float64 ----------------------------
static int legacy_solve(int io_type, int N, double *W, double *ext, double k, double n, double *r0, double *r1,
                        double tau_E, double tau_I, double dt, int max_iter, double atol,
                        double rate_soft_bound, double rate_hard_bound) {
    if (N < 1) { set_error("N must be positive"); return 1000; }
    int rc = tl_ctx.ensure_device();
    if (rc) return rc;
    const size_t dim = 2 * (size_t)N;
    cudaStream_t st = tl_ctx.stream;
    GET(0, double, dim * dim + 2 * dim, dW);                 // W | ext | r0, uploaded with one copy
    double *dE = dW + dim * dim, *dR0 = dE + dim;
    GET(3, double, dim, dR);
    GET(4, int, 2, dS);
    // stage through this thread's pinned buffer: [W | ext | r0 | result | status]
    void *pin = nullptr;
    if ((rc = tl_ctx.get_pinned((dim * dim + 3 * dim + 2) * sizeof(double), &pin))) return rc < 0 ? 1000 : rc;
    double *hW = (double *)pin, *hE = hW + dim * dim, *hR0 = hE + dim, *hR = hR0 + dim;
    int *hS = (int *)(hR + dim);
    memcpy(hW, W, dim * dim * sizeof(double));
    memcpy(hE, ext, dim * sizeof(double));
    memcpy(hR0, r0, dim * sizeof(double));
    SSN_CUDA(cudaMemcpyAsync(dW, hW, (dim * dim + 2 * dim) * sizeof(double), cudaMemcpyHostToDevice, st));
    ssn_solver sv = {};
    sv.io_type = io_type; sv.max_iter = max_iter; sv.k = k; sv.n = n;
    sv.tau_E = tau_E; sv.tau_I = tau_I; sv.dt = dt; sv.atol = atol;
    sv.rate_soft_bound = rate_soft_bound; sv.rate_hard_bound = rate_hard_bound;
    CounterLease lease;
    if ((rc = lease.acquire(st))) return rc < 0 ? 1000 : rc;
    rc = launch_fixed_point_f64(sv, 1, 1, N, dW, dE, 0, dR0, dR, dS, dS + 1, /*nonfinite_fixup=*/false, lease.ptr, st);
    lease.release();
    if (rc) return rc < 0 ? 1000 : rc;
    SSN_CUDA(cudaMemcpyAsync(hR, dR, dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    SSN_CUDA(cudaMemcpyAsync(hS, dS, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SSN_CUDA(cudaStreamSynchronize(st));
    memcpy(r0, hR, dim * sizeof(double));
    if (r1) memcpy(r1, hR, dim * sizeof(double));
    return hS[0];
}

}  // namespace ssn

using namespace ssn;

extern "C" {

int solve_dynamics_asym_power_euler(int N, double *W, double *ext, double k, double n, double *r0, double *r1,
                                    double tau_E, double tau_I, double dt, int max_iter, double atol,
                                    double rate_soft_bound, double rate_hard_bound) {
    return legacy_solve(SSN_IO_POWER, N, W, ext, k, n, r0, r1, tau_E, tau_I, dt, max_iter, atol,
                        rate_soft_bound, rate_hard_bound);
}
int solve_dynamics_asym_linear_euler(int N, double *W, double *ext, double k, double n, double *r0, double *r1,
                                     double tau_E, double tau_I, double dt, int max_iter, double atol,
                                     double rate_soft_bound, double rate_hard_bound) {
    return legacy_solve(SSN_IO_LINEAR, N, W, ext, k, n, r0, r1, tau_E, tau_I, dt, max_iter, atol,
                        rate_soft_bound, rate_hard_bound);
}
int solve_dynamics_asym_tanh_euler(int N, double *W, double *ext, double k, double n, double *r0, double *r1,
                                   double tau_E, double tau_I, double dt, int max_iter, double atol,
                                   double rate_soft_bound, double rate_hard_bound) {
    return legacy_solve(SSN_IO_TANH, N, W, ext, k, n, r0, r1, tau_E, tau_I, dt, max_iter, atol,
                        rate_soft_bound, rate_hard_bound);
}

// Scalar helpers: the host instantiation of the functions the kernels use.
double dot(int dim, const double *x, const double *y) {
    double s = 0.0;
    for (int i = 0; i < dim; ++i) s += x[i] * y[i];
    return s;
}
double rate_to_volt(double rate, double k, double n) { return pow(rate / k, 1.0 / n); }
static IoConst<double> scalar_io(int io_type, double r0, double r1, double v0, double k, double n) {
    IoConst<double> c = make_io_const<double>(io_type, k, n, r0, r1);
    c.v0 = v0;                                   // the caller's v0 wins, as in the reference signature
    c.lin_slope = k * pow(v0, n - 1.0) * n;
    c.tanh_scale = n * r0 / ((r1 - r0) * v0);
    return c;
}
double io_pow(double v, double r0, double r1, double v0, double k, double n) {
    return io_eval<double>(scalar_io(SSN_IO_POWER, r0, r1, v0, k, n), v);
}
double io_alin(double v, double r0, double r1, double v0, double k, double n) {
    return io_eval<double>(scalar_io(SSN_IO_LINEAR, r0, r1, v0, k, n), v);
}
double io_atanh(double v, double r0, double r1, double v0, double k, double n) {
    return io_eval<double>(scalar_io(SSN_IO_TANH, r0, r1, v0, k, n), v);
}

// ---- batched fixed points -----------------------------------------------------------
int ssn_fixed_point_batch(const ssn_solver *solver, int nz, int nb, int n_sites, int w_kind, const float *w,
                          const ssn_jds *jds, const float *ext, int ext_per_network, const float *r_init,
                          float *R, int *status, int *iters, int precise, int mem, void *stream) {
    int rc = validate(solver, nz, nb, n_sites);
    if (rc) return rc;
    if (nz == 0 || nb == 0) return 0;
    const size_t dim = 2 * (size_t)n_sites;
    if (mem == SSN_MEM_DEVICE) {
        cudaStream_t st = (cudaStream_t)stream;
        CounterLease lease;
        if ((rc = lease.acquire(st))) return rc;
        int *counter = lease.ptr;
        if (!precise)
            return launch_fixed_point_f32(*solver, nz, nb, n_sites, w_kind, w, jds, ext, ext_per_network, r_init,
                                          R, status, iters, counter, st);
        // float64 kernel on float32 device data: widen into scratch owned by this thread.
        if ((rc = tl_ctx.ensure_device())) return rc;
        GET(0, double, (size_t)nz * dim * dim, dW);
        GET(1, double, (size_t)(ext_per_network ? nz : 1) * nb * dim, dE);
        GET(2, double, (size_t)nz * nb * dim, dR0);
        GET(3, double, (size_t)nz * nb * dim, dR);
        GET(5, float, (size_t)nz * dim * dim, dWf);
        const float *wsrc = w;
        if (w_kind == SSN_W_FROM_Z) {
            if (!jds) { set_error("SSN_W_FROM_Z needs jds"); return -1; }
            if ((rc = launch_generate_weight(nz, n_sites, w, *jds, dWf, st))) return rc;
            wsrc = dWf;
        }
        if ((rc = launch_convert_f32_to_f64(wsrc, dW, (size_t)nz * dim * dim, st))) return rc;
        if ((rc = launch_convert_f32_to_f64(ext, dE, (size_t)(ext_per_network ? nz : 1) * nb * dim, st))) return rc;
        if (r_init && (rc = launch_convert_f32_to_f64(r_init, dR0, (size_t)nz * nb * dim, st))) return rc;
        if ((rc = launch_fixed_point_f64(*solver, nz, nb, n_sites, dW, dE, ext_per_network, r_init ? dR0 : nullptr,
                                         dR, status, iters, true, counter, st))) return rc;
        if ((rc = launch_convert_f64_to_f32(dR, R, (size_t)nz * nb * dim, st))) return rc;
        // scratch is reused by the next call of this thread: make the stream order explicit
        return check_cuda(cudaStreamSynchronize(st), "precise device path");
    }

    // host arrays: slabs of networks through two staging slots on two streams, so the H2D copy of one
    // slab and the D2H copy of the previous one overlap the solve of the current one (pinned host
    // memory makes the copies truly asynchronous; pageable memory still works, just without overlap)
    if ((rc = tl_ctx.ensure_device())) return rc;
    cudaStream_t streams[2] = {tl_ctx.stream, nullptr};
    if ((rc = tl_ctx.second_stream(&streams[1]))) return rc;
    const int slab = std::min(nz, precise ? 256 : 128);
    const size_t n_ext = (size_t)(ext_per_network ? slab : 1) * nb * dim;
    float *dw[2], *de[2], *dr0[2], *dr[2];
    int *ds[2];
    for (int q = 0; q < 2; ++q) {
        void *p;
        if ((rc = tl_ctx.get(16 + 5 * q + 0, (size_t)slab * dim * dim * sizeof(float), &p))) return rc; dw[q] = (float *)p;
        if ((rc = tl_ctx.get(16 + 5 * q + 1, n_ext * sizeof(float), &p))) return rc; de[q] = (float *)p;
        if ((rc = tl_ctx.get(16 + 5 * q + 2, (size_t)slab * nb * dim * sizeof(float), &p))) return rc; dr0[q] = (float *)p;
        if ((rc = tl_ctx.get(16 + 5 * q + 3, (size_t)slab * nb * dim * sizeof(float), &p))) return rc; dr[q] = (float *)p;
        if ((rc = tl_ctx.get(16 + 5 * q + 4, (size_t)slab * nb * 2 * sizeof(int), &p))) return rc; ds[q] = (int *)p;
        if (!ext_per_network)
            SSN_CUDA(cudaMemcpyAsync(de[q], ext, n_ext * sizeof(float), cudaMemcpyHostToDevice, streams[q]));
    }
    int q = 0;
    for (int z0 = 0; z0 < nz; z0 += slab, q ^= 1) {
        const int m = std::min(slab, nz - z0);
        cudaStream_t st = streams[q];
        SSN_CUDA(cudaMemcpyAsync(dw[q], w + (size_t)z0 * dim * dim, (size_t)m * dim * dim * sizeof(float),
                                 cudaMemcpyHostToDevice, st));
        if (ext_per_network)
            SSN_CUDA(cudaMemcpyAsync(de[q], ext + (size_t)z0 * nb * dim, (size_t)m * nb * dim * sizeof(float),
                                     cudaMemcpyHostToDevice, st));
        if (r_init)
            SSN_CUDA(cudaMemcpyAsync(dr0[q], r_init + (size_t)z0 * nb * dim, (size_t)m * nb * dim * sizeof(float),
                                     cudaMemcpyHostToDevice, st));
        rc = ssn_fixed_point_batch(solver, m, nb, n_sites, w_kind, dw[q], jds, de[q], ext_per_network,
                                   r_init ? dr0[q] : nullptr, dr[q], ds[q], ds[q] + (size_t)slab * nb, precise,
                                   SSN_MEM_DEVICE, st);
        if (rc) return rc;
        SSN_CUDA(cudaMemcpyAsync(R + (size_t)z0 * nb * dim, dr[q], (size_t)m * nb * dim * sizeof(float),
                                 cudaMemcpyDeviceToHost, st));
        SSN_CUDA(cudaMemcpyAsync(status + (size_t)z0 * nb, ds[q], (size_t)m * nb * sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
        if (iters)
            SSN_CUDA(cudaMemcpyAsync(iters + (size_t)z0 * nb, ds[q] + (size_t)slab * nb, (size_t)m * nb * sizeof(int),
                                     cudaMemcpyDeviceToHost, st));
    }
    SSN_CUDA(cudaStreamSynchronize(streams[0]));
    SSN_CUDA(cudaStreamSynchronize(streams[1]));
    return 0;
}

int ssn_fixed_point_batch_f64(const ssn_solver *solver, int nz, int nb, int n_sites, const double *W,
                              const double *ext, const double *r_init, double *R, int *status, int *iters,
                              int precise) {
    int rc = validate(solver, nz, nb, n_sites);
    if (rc) return rc;
    if (nz == 0 || nb == 0) return 0;
    if ((rc = tl_ctx.ensure_device())) return rc;
    cudaStream_t st = tl_ctx.stream;
    const size_t dim = 2 * (size_t)n_sites;
    const int slab = std::min(nz, 256);
    GET(0, double, (size_t)slab * dim * dim, dW);
    GET(1, double, (size_t)nb * dim, dE);
    GET(2, double, (size_t)slab * nb * dim, dR0);
    GET(3, double, (size_t)slab * nb * dim, dR);
    GET(4, int, (size_t)slab * nb * 2, dS);
    GET(5, float, (size_t)slab * dim * dim, fW);
    GET(6, float, (size_t)nb * dim, fE);
    GET(7, float, (size_t)slab * nb * dim * 2, fR);
    SSN_CUDA(cudaMemcpyAsync(dE, ext, (size_t)nb * dim * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!precise && (rc = launch_convert_f64_to_f32(dE, fE, (size_t)nb * dim, st))) return rc;
    CounterLease lease;
    int *counter = nullptr;
    for (int z0 = 0; z0 < nz; z0 += slab) {
        const int m = std::min(slab, nz - z0);
        const size_t nR = (size_t)m * nb * dim;
        SSN_CUDA(cudaMemcpyAsync(dW, W + (size_t)z0 * dim * dim, (size_t)m * dim * dim * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
        if (r_init)
            SSN_CUDA(cudaMemcpyAsync(dR0, r_init + (size_t)z0 * nb * dim, nR * sizeof(double),
                                     cudaMemcpyHostToDevice, st));
        if (precise) {
            if ((rc = lease.acquire(st))) return rc;
            counter = lease.ptr;
            rc = launch_fixed_point_f64(*solver, m, nb, n_sites, dW, dE, 0, r_init ? dR0 : nullptr, dR, dS,
                                        dS + (size_t)slab * nb, true, counter, st);
            if (rc) return rc;
        } else {
            float *fR0 = fR + (size_t)slab * nb * dim;
            if ((rc = launch_convert_f64_to_f32(dW, fW, (size_t)m * dim * dim, st))) return rc;
            if (r_init && (rc = launch_convert_f64_to_f32(dR0, fR0, nR, st))) return rc;
            if ((rc = lease.acquire(st))) return rc;
            counter = lease.ptr;
            rc = launch_fixed_point_f32(*solver, m, nb, n_sites, SSN_W_DENSE, fW, nullptr, fE, 0,
                                        r_init ? fR0 : nullptr, fR, dS, dS + (size_t)slab * nb, counter, st);
            if (rc) return rc;
            if ((rc = launch_convert_f32_to_f64(fR, dR, nR, st))) return rc;
        }
        SSN_CUDA(cudaMemcpyAsync(R + (size_t)z0 * nb * dim, dR, nR * sizeof(double), cudaMemcpyDeviceToHost, st));
        SSN_CUDA(cudaMemcpyAsync(status + (size_t)z0 * nb, dS, (size_t)m * nb * sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
        if (iters)
            SSN_CUDA(cudaMemcpyAsync(iters + (size_t)z0 * nb, dS + (size_t)slab * nb, (size_t)m * nb * sizeof(int),
                                     cudaMemcpyDeviceToHost, st));
        SSN_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---- streamed host path: what ssnode.find_fixed_points calls ------------------------------------
// The reference hands the solver one float64 W (1.29 MB at 2N = 402) per network, each its own numpy array
// (tc_gan/ssnode.py:436-447).  Here the caller passes the list of pointers as is: no concatenation on the Python
// side.  Host threads round slab k+1 to float32 straight into pinned staging while the GPU solves slab k; three
// pinned slabs rotate over two streams, so H2D, solve and D2H of neighbouring slabs overlap.
}  // extern "C"
namespace {

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (cap >= bytes) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        SSN_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
        cap = bytes;
        return 0;
    }
};
struct StreamCtx {
    int device = -1;
    static constexpr int NBUF = 3;
    PinnedBuf win[NBUF], rout[NBUF];
    void *dwin[NBUF] = {nullptr, nullptr, nullptr}, *drout[NBUF] = {nullptr, nullptr, nullptr};
    size_t dwin_cap[NBUF] = {0, 0, 0}, drout_cap[NBUF] = {0, 0, 0};
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t h2d_done[NBUF] = {nullptr, nullptr, nullptr}, out_done[NBUF] = {nullptr, nullptr, nullptr};
    float *dext = nullptr;
    size_t dext_cap = 0;
    int prepare() {
        int dev = 0;
        SSN_CUDA(cudaGetDevice(&dev));
        if (dev == device) return 0;
        device = dev;                                   // (buffers of a previous device are abandoned, not freed)
        for (int q = 0; q < NBUF; ++q) { win[q] = PinnedBuf(); rout[q] = PinnedBuf(); dwin[q] = drout[q] = nullptr; dwin_cap[q] = drout_cap[q] = 0; }
        dext = nullptr; dext_cap = 0;
        for (int q = 0; q < 2; ++q) SSN_CUDA(cudaStreamCreateWithFlags(&stream[q], cudaStreamNonBlocking));
        for (int q = 0; q < NBUF; ++q) {
            SSN_CUDA(cudaEventCreateWithFlags(&h2d_done[q], cudaEventDisableTiming));
            SSN_CUDA(cudaEventCreateWithFlags(&out_done[q], cudaEventDisableTiming));
        }
        return 0;
    }
    static int grow(void **p, size_t *cap, size_t bytes) {
        if (*cap >= bytes) return 0;
        if (*p) cudaFree(*p);
        *p = nullptr; *cap = 0;
        SSN_CUDA(cudaMalloc(p, bytes));
        *cap = bytes;
        return 0;
    }
};
static std::mutex g_stream_ctx_mutex;
static std::vector<StreamCtx *> g_stream_ctx_free;
struct StreamCtxLease {
    StreamCtx *p = nullptr;
    StreamCtxLease() {
        std::lock_guard<std::mutex> lock(g_stream_ctx_mutex);
        if (!g_stream_ctx_free.empty()) { p = g_stream_ctx_free.back(); g_stream_ctx_free.pop_back(); }
        else p = new StreamCtx();
    }
    ~StreamCtxLease() {
        std::lock_guard<std::mutex> lock(g_stream_ctx_mutex);
        g_stream_ctx_free.push_back(p);
    }
};

// run fn(i) for i in [0, n) on up to `threads` host threads (the calling thread included)
template <class F>
static void parallel_for(int n, int threads, F fn) {
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::atomic<int> next{0};
    auto body = [&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); };
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (int t = 1; t < threads; ++t) pool.emplace_back(body);
    body();
    for (auto &th : pool) th.join();
}

}  // namespace
extern "C" {

// dst[i] = copy of the `bytes` bytes at items[i], i < n, with several host threads (np.array() of a list of
// per-network arrays is a single-threaded 1.3 GB copy at the benchmark size)
int ssn_host_gather(const void *const *items, int n, size_t bytes, void *dst, int host_threads) {
    if (n < 0 || (n > 0 && (!items || !dst))) { set_error("ssn_host_gather: bad arguments"); return -1; }
    if (host_threads <= 0) host_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    parallel_for(n, host_threads, [&](int i) { memcpy((char *)dst + (size_t)i * bytes, items[i], bytes); });
    return 0;
}

int ssn_fixed_point_batch_ptrs(const ssn_solver *solver, int nz, int nb, int n_sites, int w_kind,
                               const void *const *items, int items_f32, const ssn_jds *jds, const double *ext,
                               const double *r_init, double *R, int *status, int *iters, int host_threads) {
    int rc = validate(solver, nz, nb, n_sites);
    if (rc) return rc;
    if (nz == 0 || nb == 0) return 0;
    if (!items || !ext || !R || !status) { set_error("ssn_fixed_point_batch_ptrs: NULL argument"); return -1; }
    if (w_kind == SSN_W_FROM_Z && !jds) { set_error("SSN_W_FROM_Z needs jds"); return -1; }
    StreamCtxLease lease;
    StreamCtx &c = *lease.p;
    if ((rc = c.prepare())) return rc;
    const size_t dim = 2 * (size_t)n_sites, wsz = dim * dim, rsz = (size_t)nb * dim;
    // slab: large enough that the persistent kernel's tail (one wave of its ~15-22 resident clusters) is a small
    // part of a launch, small enough that the first conversion and the last solve, which nothing overlaps, stay short
    const int slab = std::min(nz, std::max(32, std::min(256, (int)((192u << 20) / (wsz * sizeof(float))))));
    if (host_threads <= 0) host_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    const size_t out_ints = 2 * (size_t)nb;                               // status + iters per network
    const size_t rin = r_init ? rsz : 0;
    for (int q = 0; q < StreamCtx::NBUF; ++q) {
        if ((rc = c.win[q].ensure((size_t)slab * (wsz + rin) * sizeof(float)))) return rc;
        if ((rc = c.rout[q].ensure((size_t)slab * (rsz * sizeof(float) + out_ints * sizeof(int))))) return rc;
        if ((rc = StreamCtx::grow(&c.dwin[q], &c.dwin_cap[q], (size_t)slab * (wsz + rin) * sizeof(float)))) return rc;
        if ((rc = StreamCtx::grow(&c.drout[q], &c.drout_cap[q], (size_t)slab * (rsz * sizeof(float) + out_ints * sizeof(int))))) return rc;
    }
    if ((rc = StreamCtx::grow((void **)&c.dext, &c.dext_cap, rsz * sizeof(float)))) return rc;
    {
        std::vector<float> e32(rsz);
        for (size_t i = 0; i < rsz; ++i) e32[i] = (float)ext[i];
        SSN_CUDA(cudaMemcpy(c.dext, e32.data(), rsz * sizeof(float), cudaMemcpyHostToDevice));
    }
    auto drain = [&](int q, int z0, int m) {                             // results of the slab that used buffer q
        if (cudaEventSynchronize(c.out_done[q]) != cudaSuccess) return check_cuda(cudaGetLastError(), "streamed results");
        const float *hr = (const float *)c.rout[q].p;
        const int *hs = (const int *)(hr + (size_t)slab * rsz);
        parallel_for(m, host_threads, [&](int i) {
            double *dst = R + (size_t)(z0 + i) * rsz;
            const float *src = hr + (size_t)i * rsz;
            for (size_t e = 0; e < rsz; ++e) dst[e] = (double)src[e];
        });
        memcpy(status + (size_t)z0 * nb, hs, (size_t)m * nb * sizeof(int));
        if (iters) memcpy(iters + (size_t)z0 * nb, hs + (size_t)slab * nb, (size_t)m * nb * sizeof(int));
        return 0;
    };
    struct Pending { int q, z0, m; };
    std::vector<Pending> pending;
    int k = 0;
    for (int z0 = 0; z0 < nz; z0 += slab, ++k) {
        const int m = std::min(slab, nz - z0), q = k % StreamCtx::NBUF;
        cudaStream_t st = c.stream[k & 1];
        if (k >= StreamCtx::NBUF) {                                       // buffer q is about to be reused
            const Pending p = pending.front();
            pending.erase(pending.begin());
            if ((rc = drain(p.q, p.z0, p.m))) return rc;                  // (out_done[q] also implies h2d_done[q])
        }
        float *hw = (float *)c.win[q].p;
        float *hr0 = hw + (size_t)slab * wsz;
        parallel_for(m, host_threads, [&](int i) {
            float *dst = hw + (size_t)i * wsz;
            if (items_f32) {
                memcpy(dst, items[z0 + i], wsz * sizeof(float));
            } else {
                const double *src = (const double *)items[z0 + i];
                for (size_t e = 0; e < wsz; ++e) dst[e] = (float)src[e];
            }
            if (r_init) {
                const double *r0 = r_init + (size_t)(z0 + i) * rsz;
                for (size_t e = 0; e < rsz; ++e) hr0[(size_t)i * rsz + e] = (float)r0[e];
            }
        });
        float *dw = (float *)c.dwin[q], *dr0 = dw + (size_t)slab * wsz;
        float *dr = (float *)c.drout[q];
        int *ds = (int *)(dr + (size_t)slab * rsz);
        SSN_CUDA(cudaMemcpyAsync(dw, hw, (size_t)m * wsz * sizeof(float), cudaMemcpyHostToDevice, st));
        if (r_init) SSN_CUDA(cudaMemcpyAsync(dr0, hr0, (size_t)m * rsz * sizeof(float), cudaMemcpyHostToDevice, st));
        SSN_CUDA(cudaEventRecord(c.h2d_done[q], st));
        {
            CounterLease cl;
            if ((rc = cl.acquire(st))) return rc;
            rc = launch_fixed_point_f32(*solver, m, nb, n_sites, w_kind, dw, jds, c.dext, 0, r_init ? dr0 : nullptr,
                                        dr, ds, ds + (size_t)slab * nb, cl.ptr, st);
            if (rc) return rc;
        }
        float *hr = (float *)c.rout[q].p;
        int *hs = (int *)(hr + (size_t)slab * rsz);
        SSN_CUDA(cudaMemcpyAsync(hr, dr, (size_t)m * rsz * sizeof(float), cudaMemcpyDeviceToHost, st));
        SSN_CUDA(cudaMemcpyAsync(hs, ds, (size_t)m * nb * sizeof(int), cudaMemcpyDeviceToHost, st));
        SSN_CUDA(cudaMemcpyAsync(hs + (size_t)slab * nb, ds + (size_t)slab * nb, (size_t)m * nb * sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
        SSN_CUDA(cudaEventRecord(c.out_done[q], st));
        pending.push_back({q, z0, m});
    }
    for (const Pending &p : pending)
        if ((rc = drain(p.q, p.z0, p.m))) return rc;
    return 0;
}

int ssn_generate_weight(int nz, int n_sites, const float *z, const ssn_jds *jds, float *W, int mem, void *stream) {
    if (!jds || nz < 0 || n_sites < 1) { set_error("bad arguments"); return -1; }
    if (mem == SSN_MEM_DEVICE) return launch_generate_weight(nz, n_sites, z, *jds, W, (cudaStream_t)stream);
    int rc = tl_ctx.ensure_device();
    if (rc) return rc;
    const size_t total = (size_t)nz * 4 * n_sites * n_sites;
    GET(8, float, total, dz);
    GET(5, float, total, dW);
    cudaStream_t st = tl_ctx.stream;
    SSN_CUDA(cudaMemcpyAsync(dz, z, total * sizeof(float), cudaMemcpyHostToDevice, st));
    if ((rc = launch_generate_weight(nz, n_sites, dz, *jds, dW, st))) return rc;
    SSN_CUDA(cudaMemcpyAsync(W, dW, total * sizeof(float), cudaMemcpyDeviceToHost, st));
    return check_cuda(cudaStreamSynchronize(st), "generate_weight");
}

// ---- gradients ------------------------------------------------------------------------
int ssn_ift_gradient_batch(const ssn_solver *solver, int nz, int nb, int n_sites, const float *z,
                           const ssn_jds *jds, const float *ext, int ext_per_network, const float *R,
                           const float *g, double rtol, double *grad, float *mu, int *status, int *iters,
                           float *grad_ext, int mem, void *stream) {
    int rc = validate(solver, nz, nb, n_sites);
    if (rc) return rc;
    if (!jds) { set_error("jds is NULL"); return -1; }
    const size_t dim = 2 * (size_t)n_sites;
    CounterLease lease;
    if (mem == SSN_MEM_DEVICE) {
        if ((rc = lease.acquire((cudaStream_t)stream))) return rc;
        return launch_ift_gradient(*solver, nz, nb, n_sites, z, *jds, ext, ext_per_network, R, g, rtol, grad, mu,
                                   status, iters, grad_ext, lease.ptr, (cudaStream_t)stream);
    }
    if ((rc = tl_ctx.ensure_device())) return rc;
    cudaStream_t st = tl_ctx.stream;
    if ((rc = lease.acquire(st))) return rc;
    int *counter = lease.ptr;
    const size_t nR = (size_t)nz * nb * dim, nE = (size_t)(ext_per_network ? nz : 1) * nb * dim;
    GET(8, float, (size_t)nz * dim * dim, dz);
    GET(9, float, nE, de);
    GET(10, float, nR, dr);
    GET(11, float, nR, dg);
    GET(12, int, (size_t)nz * nb * 2, ds);
    GET(13, float, nR, dmu);
    GET(14, double, 12, dgrad);
    GET(15, float, nR, dge);
    SSN_CUDA(cudaMemcpyAsync(dz, z, (size_t)nz * dim * dim * sizeof(float), cudaMemcpyHostToDevice, st));
    SSN_CUDA(cudaMemcpyAsync(de, ext, nE * sizeof(float), cudaMemcpyHostToDevice, st));
    SSN_CUDA(cudaMemcpyAsync(dr, R, nR * sizeof(float), cudaMemcpyHostToDevice, st));
    SSN_CUDA(cudaMemcpyAsync(dg, g, nR * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = launch_ift_gradient(*solver, nz, nb, n_sites, dz, *jds, de, ext_per_network, dr, dg, rtol, dgrad, dmu, ds,
                             ds + (size_t)nz * nb, grad_ext ? dge : nullptr, counter, st);
    if (rc) return rc;
    SSN_CUDA(cudaMemcpyAsync(grad, dgrad, 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (mu) SSN_CUDA(cudaMemcpyAsync(mu, dmu, nR * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (grad_ext) SSN_CUDA(cudaMemcpyAsync(grad_ext, dge, nR * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (status) SSN_CUDA(cudaMemcpyAsync(status, ds, (size_t)nz * nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (iters)
        SSN_CUDA(cudaMemcpyAsync(iters, ds + (size_t)nz * nb, (size_t)nz * nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    return check_cuda(cudaStreamSynchronize(st), "ift gradient");
}

int ssn_euler_forward(const ssn_solver *solver, int nz, int nb, int n_sites, const float *z, const ssn_jds *jds,
                      const float *ext, int ext_per_network, int seqlen, int skip_steps,
                      double rate_penalty_threshold, float *time_avg, double *penalties, float *traj, float *gain,
                      void *stream) {
    int rc = validate(solver, nz, nb, n_sites);
    if (rc) return rc;
    if (!jds || seqlen < 1 || skip_steps < 0 || skip_steps >= seqlen) {
        set_error("bad euler arguments (seqlen=%d skip_steps=%d)", seqlen, skip_steps);
        return -1;
    }
    CounterLease lease;
    if ((rc = lease.acquire((cudaStream_t)stream))) return rc;
    return launch_euler_forward(*solver, nz, nb, n_sites, z, *jds, ext, ext_per_network, seqlen, skip_steps,
                                rate_penalty_threshold, time_avg, penalties, traj, gain, lease.ptr,
                                (cudaStream_t)stream);
}

int ssn_euler_backward(const ssn_solver *solver, int nz, int nb, int n_sites, const float *z, const ssn_jds *jds,
                       int seqlen, int skip_steps, double rate_penalty_threshold, const float *grad_time_avg,
                       double w_dyn, double w_rate, const float *w_dev, const float *traj, const float *gain,
                       float *adj, double *grad, float *grad_ext, void *stream) {
    int rc = validate(solver, nz, nb, n_sites);
    if (rc) return rc;
    if (!jds || !traj || !gain || !adj) { set_error("euler backward needs jds, traj, gain, adj"); return -1; }
    CounterLease lease;
    if ((rc = lease.acquire((cudaStream_t)stream))) return rc;
    return launch_euler_backward(*solver, nz, nb, n_sites, z, *jds, seqlen, skip_steps, rate_penalty_threshold,
                                 grad_time_avg, w_dyn, w_rate, w_dev, traj, gain, adj, grad, grad_ext, lease.ptr,
                                 (cudaStream_t)stream);
}

// The parameter-gradient contraction alone (development / tests): grad[12] = < sum_k adj_k traj_k^T, dW/dtheta >.
int ssn_bptt_param_grad(int nz, int nb, int n_sites, int seqlen, const float *adj, const float *traj, const float *z,
                        const ssn_jds *jds, double *grad, void *stream) {
    if (!adj || !traj || !z || !jds || !grad || nz < 0 || nb < 1 || seqlen < 1 || n_sites < 1) {
        set_error("ssn_bptt_param_grad: bad arguments");
        return -1;
    }
    return launch_bptt_param_grad(nz, nb, n_sites, seqlen, adj, traj, z, *jds, grad, (cudaStream_t)stream);
}

// ---- probes (output-side boundary) ---------------------------------------------------------
int ssn_probe_gather(const float *rates, const int *model_ids, const int *probes, int batch, int nz, int nb,
                     int n_sites, float *out, void *stream) {
    if (!rates || !model_ids || !probes || !out || batch < 0 || nz < 0 || nb < 0 || n_sites < 1) {
        set_error("ssn_probe_gather: bad arguments");
        return -1;
    }
    return launch_probe_gather(rates, model_ids, probes, batch, nz, nb, 2 * n_sites, out, (cudaStream_t)stream);
}
int ssn_probe_scatter(const float *grad_out, const int *model_ids, const int *probes, int batch, int nz, int nb,
                      int n_sites, float *grad_rates, void *stream) {
    if (!grad_out || !model_ids || !probes || !grad_rates || batch < 0 || nz < 0 || nb < 0 || n_sites < 1) {
        set_error("ssn_probe_scatter: bad arguments");
        return -1;
    }
    return launch_probe_scatter(grad_out, model_ids, probes, batch, nz, nb, 2 * n_sites, grad_rates, (cudaStream_t)stream);
}

// ---- introspection ----------------------------------------------------------------------
int ssn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char *ssn_last_error(void) { return tl_error; }
int ssn_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
int ssn_fixed_point_occupancy(int n_sites, int *cluster_size, int *resident_clusters) {
    return fixed_point_occupancy(n_sites, cluster_size, resident_clusters);
}
int ssn_traj_pitch(int n_sites) { return traj_pitch(n_sites); }
int ssn_fixed_point_kernel_name(int n_sites, char *buf, int cap) {
    return fixed_point_kernel_name(n_sites, buf, cap);
}
int ssn_profile_enable(int on) {
    const int was = g_profile.exchange(on ? 1 : 0);
    if (!on) {
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        for (auto &t : g_timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
        g_timed.clear();
    }
    return was;
}
// "name total_ms launches\n" per kernel since the last read; waits for the recorded launches to finish.
int ssn_profile_read(char *buf, int cap) {
    std::vector<TimedLaunch> taken;
    {
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        taken.swap(g_timed);
    }
    struct Acc { const char *name; double ms; int n; };
    std::vector<Acc> acc;
    for (auto &t : taken) {
        float ms = 0.f;
        if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
            size_t i = 0;
            for (; i < acc.size(); ++i) if (!strcmp(acc[i].name, t.name)) break;
            if (i == acc.size()) acc.push_back({t.name, 0.0, 0});
            acc[i].ms += ms; acc[i].n += 1;
        } else {
            cudaGetLastError();
        }
        cudaEventDestroy(t.e0); cudaEventDestroy(t.e1);
    }
    int used = 0;
    if (buf && cap > 0) buf[0] = 0;
    for (auto &a : acc) {
        const int w = snprintf(buf ? buf + used : nullptr, buf && cap > used ? cap - used : 0, "%s %.6f %d\n", a.name, a.ms, a.n);
        if (w < 0 || used + w >= cap) break;
        used += w;
    }
    return (int)acc.size();
}
int ssn_measure_fp32_peak(double *tflops) {
    int dev = 0, sms = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *d = nullptr;
    SSN_CUDA(cudaMalloc(&d, sizeof(float)));
    cudaEvent_t e0, e1;
    SSN_CUDA(cudaEventCreate(&e0));
    SSN_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 2;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        SSN_CUDA(cudaEventRecord(e0, 0));
        ffma_peak_kernel<<<blocks, 1024>>>(d, iters, 0.999f, 1e-4f);
        SSN_CUDA(cudaEventRecord(e1, 0));
        SSN_CUDA(cudaEventSynchronize(e1));
        count_launch();
        float ms = 0.f;
        SSN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * 1024.0 * blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (tflops) *tflops = best;
    return 0;
}

}  // extern "C"
