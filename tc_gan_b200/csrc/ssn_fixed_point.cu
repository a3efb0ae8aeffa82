// K1: batched SSN fixed-point solve (replaces tc_gan/ext/ssnode.c:69-187 driven by the
// thread pool of tc_gan/ssnode.py:423-510, and tc_gan/weight_gen.py:13-26).
//
//   r <- r + (dt/tau) (-r + f(W r + I)),  Jacobi sweeps until max|dr| < atol.
//
// ssn_fp_cluster_kernel: persistent thread-block clusters pull networks from a
// global work counter.  W (or W built from z) stays in the cluster's shared
// memory; the 8-stimulus state panel is exchanged through DSMEM once per sweep;
// the O(2N) state update and the convergence test run in float64, the O((2N)^2)
// contraction in FP32 FFMA.
//
// ssn_fp64_kernel: all-float64 variant (one CTA per network x 8-stimulus panel,
// W streamed from L2) used by the reference-ABI single-solve symbols and by
// `precise` batched calls.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include "ssn_cluster_core.cuh"
#include <cstdio>
#include <cstring>
#include "ssn_launch.h"

namespace ssn {

struct FpArgs {
    int nz, nb, n_sites;
    ClusterShape shape;
    int w_kind;
    const float *w;
    WeightConst wc;
    const float *ext;
    long long ext_stride_z;        // 0: one [nb][dim] table shared by all networks
    const float *r_init;           // [nz][nb][dim] or null
    float *R;
    int *status, *iters;
    int *work_counter;
    IoConst<float> io;
    double eps_E, eps_I, atol, r_hard;
    int max_iter, check_hard;
};

template <int TI, int KL, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) ssn_fp_cluster_kernel(const FpArgs a) {
    using Own = Owner<TI, KL>;
    constexpr int TO = Own::TO;
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int csize = a.shape.csize;
    const int dim = a.shape.dim, kpad = a.shape.kpad, rpc = a.shape.rpc;
    const int P = panel_P(kpad);
    const SmemLayout L = smem_layout(a.shape, a.n_sites);
    float *Wsm = reinterpret_cast<float *>(smem + L.w_off);
    float *Xf = reinterpret_cast<float *>(smem + L.x_off);
    const float4 *X4 = reinterpret_cast<const float4 *>(smem + L.x_off);
    float *extsm = reinterpret_cast<float *>(smem + L.ext_off);
    float *gtab = reinterpret_cast<float *>(smem + L.gtab_off);
    Misc *misc = reinterpret_cast<Misc *>(smem + L.misc_off);

    const int tid = threadIdx.x, nthreads = NWARPS * 32;
    const int warp = tid >> 5, lane = tid & 31;
    const int kl = lane % KL;
    const int grp = warp * (32 / KL) + lane / KL;
    const int row_base = rank * rpc;
    const int rows_here = max(0, min(rpc, dim - row_base));

    int wrow[TI];                                            // smem rows this thread contracts
#pragma unroll
    for (int t = 0; t < TI; ++t) wrow[t] = min(grp * TI + t, rows_here - 1);

    // rows / stimulus this thread owns after the reduce-scatter
    const int my_stim = Own::stim(kl);
    const int own0 = grp * TI + Own::first_row(kl);          // first owned local row
    bool valid[TO];
#pragma unroll
    for (int u = 0; u < TO; ++u)
        valid[u] = (Own::first_row(kl) + u < TI) && (own0 + u < rows_here);

    // peers' panels and flag words through distributed shared memory
    unsigned xpeer[MAX_CLUSTER];
#pragma unroll
    for (int p = 0; p < MAX_CLUSTER; ++p) xpeer[p] = map_to_rank(smem_u32(Xf), p < csize ? p : 0);

    if (a.w_kind == SSN_W_FROM_Z) build_profile_table(a.wc, a.n_sites, gtab, tid, nthreads);
    for (int i = tid; i < 2 * 2 * P * 4; i += nthreads) Xf[i] = 0.f;     // padding columns stay 0
    if (tid == 0) misc->myflags = 0u;
    __syncthreads();

    const int n_chunks = (a.nb + TB - 1) / TB;
    const float atol_f = (float)a.atol;
    (void)atol_f;

    for (;;) {
        // ---- next network from the global queue ----
        if (rank == 0 && tid == 0) {
            const int n = atomicAdd(a.work_counter, 1);
            for (int p = 0; p < csize; ++p) st_cluster_u32(map_to_rank(smem_u32(&misc->next_net), p), (unsigned)n);
        }
        cluster.sync();
        const int net = misc->next_net;
        if (net >= a.nz) break;

        load_matrix_slice<false>(Wsm, a.w + (size_t)net * dim * dim, a.w_kind, a.wc, gtab,
                                 a.n_sites, dim, kpad, row_base, rows_here, tid, nthreads);

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int b0 = chunk * TB;
            const int nact = min(TB, a.nb - b0);
            const bool active = my_stim < nact;
            const float *ext_net = a.ext + (size_t)net * a.ext_stride_z;

            // stimulus rows of this CTA -> smem, [local row][8]
            for (int i = tid; i < rows_here * TB; i += nthreads) {
                const int r = i / TB, b = i % TB;
                extsm[i] = b < nact ? __ldg(ext_net + (size_t)(b0 + b) * dim + row_base + r) : 0.f;
            }

            // initial state (float64 master copy in registers), published to every CTA
            double rstate[TO];
            float ext_own[TO];
            double eps_own[TO];
            unsigned xoff[TO];                                // byte offset of (row, stim) in panel 0
#pragma unroll
            for (int u = 0; u < TO; ++u) {
                rstate[u] = 0.0;
                ext_own[u] = 0.f;
                eps_own[u] = 0.0;
                xoff[u] = 0u;
                if (valid[u]) {
                    const int gr = row_base + own0 + u;
                    if (active && a.r_init)
                        rstate[u] = (double)__ldg(a.r_init + ((size_t)net * a.nb + b0 + my_stim) * dim + gr);
                    eps_own[u] = gr < a.n_sites ? a.eps_E : a.eps_I;
                    xoff[u] = 4u * (unsigned)panel_index(P, 0, gr, my_stim);
                    const float rf = (float)rstate[u];
#pragma unroll
                    for (int p = 0; p < MAX_CLUSTER; ++p)
                        if (p < csize) st_cluster_f32(xpeer[p] + xoff[u], rf);
                }
            }
            cluster.sync();
#pragma unroll
            for (int u = 0; u < TO; ++u)
                if (valid[u]) ext_own[u] = extsm[(own0 + u) * TB + my_stim];

            unsigned done = nact >= TB ? 0u : (0xffu << nact) & 0xffu;   // finished stimuli
            int my_status = 1, my_iters = a.max_iter;                   // of stimulus my_stim
            int buf = 0;
            const unsigned buf_bytes = 2u * (unsigned)P * 16u;           // one panel buffer (2 planes)

            for (int it = 1; it <= a.max_iter; ++it) {
                float acc[TI][TB], v[TO];
                contract_panel<TI, KL>(acc, Wsm, X4, P, kpad, buf, wrow, kl);
                reduce_scatter<TI, KL>(acc, v, kl);

                const int nbuf = buf ^ 1;
                const bool frozen = (done >> my_stim) & 1u;
                bool moving = false, above = false;
#pragma unroll
                for (int u = 0; u < TO; ++u) {
                    if (valid[u]) {
                        const float fv = io_eval<float>(a.io, v[u] + ext_own[u]);
                        const double r_old = rstate[u];
                        const double r_new = r_old + ((double)fv - r_old) * eps_own[u];
                        if (!frozen) {
                            moving |= fabs(r_new - r_old) >= a.atol;
                            above |= r_new >= a.r_hard;
                            rstate[u] = r_new;
                        }
                        const float rf = (float)rstate[u];
                        const unsigned off = xoff[u] + (nbuf ? buf_bytes : 0u);
#pragma unroll
                        for (int p = 0; p < MAX_CLUSTER; ++p)
                            if (p < csize) st_cluster_f32(xpeer[p] + off, rf);
                    }
                }
                const unsigned mm = stim_mask<KL>(moving), ma = stim_mask<KL>(above);
                if (lane == 0 && (mm | ma)) atomicOr(&misc->myflags, mm | (ma << 8));
                __syncthreads();
                if (tid == 0) {
                    const unsigned f = misc->myflags;
                    misc->myflags = 0u;
                    for (int p = 0; p < csize; ++p)
                        st_cluster_u32(map_to_rank(smem_u32(&misc->flags[nbuf][rank]), p), f);
                }
                cluster.sync();
                unsigned F = 0u;
                for (int p = 0; p < csize; ++p) F |= misc->flags[nbuf][p];
                const unsigned moving_all = F & 0xffu, above_all = (F >> 8) & 0xffu;
                // reference order: convergence first, then the hard bound (ssnode.c:84-102)
                const unsigned conv_now = ~moving_all & ~done & 0xffu;
                const unsigned hard_now = a.check_hard ? (above_all & ~done & ~conv_now & 0xffu) : 0u;
                if ((conv_now >> my_stim) & 1u) { my_status = 0; my_iters = it; }
                if ((hard_now >> my_stim) & 1u) { my_status = 2; my_iters = it; }
                done |= conv_now | hard_now;
                buf = nbuf;
                if (done == 0xffu) break;
            }

            // ---- results ----
#pragma unroll
            for (int u = 0; u < TO; ++u)
                if (valid[u] && active)
                    a.R[((size_t)net * a.nb + b0 + my_stim) * dim + row_base + own0 + u] = (float)rstate[u];
            // warp 0, first row group: one lane per stimulus (the lane with sub-index 0)
            if (rank == 0 && tid < KL && (kl % Own::SPLIT) == 0 && active) {
                a.status[(size_t)net * a.nb + b0 + my_stim] = my_status;
                if (a.iters) a.iters[(size_t)net * a.nb + b0 + my_stim] = my_iters;
            }
            // every CTA must be out of the sweep loop before the panel is re-initialised
            cluster.sync();
        }
    }
}

// A converged solve with a non-finite state counts as failed (tc_gan/ssnode.py:257-262).
__global__ void ssn_status_fixup_kernel(const float *R, int *status, int n_solves, int dim) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_solves) return;
    bool bad = false;
    for (int i = lane; i < dim; i += 32) bad |= !isfinite(R[(size_t)warp * dim + i]);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0 && bad && status[warp] == 0) status[warp] = 1;
}

// ------------------------------------------------------------------------------------
// float64 kernel: one CTA per (network, panel of TBD stimuli); W (double) read from
// global memory (L2-resident) every sweep, one warp per row, lanes stride the columns.
// ------------------------------------------------------------------------------------
struct Fp64Args {
    int nz, nb, n_sites, dim;
    const double *W;               // [nz][dim][dim]
    const double *ext;             // [nb][dim] or [nz][nb][dim]
    long long ext_stride_z;
    const double *r_init;          // [nz][nb][dim] or null
    double *R;                     // [nz][nb][dim]
    int *status, *iters;
    IoConst<double> io;
    double eps_E, eps_I, atol, r_hard;
    int max_iter, check_hard;
};

template <int TBD>
__global__ void __launch_bounds__(512) ssn_fp64_kernel(const Fp64Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int dim = a.dim;
    double *X = reinterpret_cast<double *>(smem);            // [2][dim][TBD]
    double *E = X + 2 * dim * TBD;                           // [dim][TBD]
    __shared__ unsigned s_flags[2];
    const int n_chunks = (a.nb + TBD - 1) / TBD;
    const int net = blockIdx.x / n_chunks, chunk = blockIdx.x % n_chunks;
    const int b0 = chunk * TBD, nact = min(TBD, a.nb - b0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const double *W = a.W + (size_t)net * dim * dim;
    const double *ext = a.ext + (size_t)net * a.ext_stride_z;

    for (int i = tid; i < dim * TBD; i += blockDim.x) {
        const int r = i / TBD, b = i % TBD;
        E[i] = b < nact ? ext[(size_t)(b0 + b) * dim + r] : 0.0;
        X[i] = (b < nact && a.r_init) ? a.r_init[((size_t)net * a.nb + b0 + b) * dim + r] : 0.0;
    }
    if (tid < 2) s_flags[tid] = 0u;
    __syncthreads();

    unsigned done = nact >= TBD ? 0u : (((1u << TBD) - 1u) & ~((1u << nact) - 1u));
    const unsigned all = (1u << TBD) - 1u;
    int buf = 0;
    int st[TBD], its[TBD];
#pragma unroll
    for (int b = 0; b < TBD; ++b) { st[b] = 1; its[b] = a.max_iter; }

    for (int it = 1; it <= a.max_iter; ++it) {
        const double *Xc = X + buf * dim * TBD;
        double *Xn = X + (buf ^ 1) * dim * TBD;
        unsigned moving = 0u, above = 0u;
        for (int i = warp; i < dim; i += nwarps) {
            double acc[TBD];
#pragma unroll
            for (int b = 0; b < TBD; ++b) acc[b] = 0.0;
            const double *row = W + (size_t)i * dim;
            for (int j = lane; j < dim; j += 32) {
                const double w = row[j];
#pragma unroll
                for (int b = 0; b < TBD; ++b) acc[b] = fma(w, Xc[j * TBD + b], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < TBD; ++b)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
            if (lane < TBD) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < TBD; ++b) s = lane == b ? acc[b] : s;
                const int b = lane;
                const double r_old = Xc[i * TBD + b];
                double r_new = r_old;
                if (!((done >> b) & 1u)) {
                    const double fv = io_eval<double>(a.io, s + E[i * TBD + b]);
                    const double eps = i < a.n_sites ? a.eps_E : a.eps_I;
                    r_new = r_old + (fv - r_old) * eps;
                    if (fabs(r_new - r_old) >= a.atol) moving |= 1u << b;
                    if (r_new >= a.r_hard) above |= 1u << b;
                }
                Xn[i * TBD + b] = r_new;
            }
        }
        if (moving | above) atomicOr(&s_flags[it & 1], moving | (above << 8));
        __syncthreads();
        const unsigned F = s_flags[it & 1];
        if (tid == 0) s_flags[(it + 1) & 1] = 0u;
        const unsigned conv_now = ~(F & 0xffu) & ~done & all;
        const unsigned hard_now = a.check_hard ? ((F >> 8) & ~done & ~conv_now & all) : 0u;
#pragma unroll
        for (int b = 0; b < TBD; ++b) {
            if ((conv_now >> b) & 1u) { st[b] = 0; its[b] = it; }
            if ((hard_now >> b) & 1u) { st[b] = 2; its[b] = it; }
        }
        done |= conv_now | hard_now;
        buf ^= 1;
        if (done == all) break;
        __syncthreads();      // s_flags reset visible before the next sweep's atomics
    }
    __syncthreads();
    const double *Xc = X + buf * dim * TBD;
    for (int i = tid; i < dim * nact; i += blockDim.x) {
        const int b = i / dim, r = i % dim;
        a.R[((size_t)net * a.nb + b0 + b) * dim + r] = Xc[r * TBD + b];
    }
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < TBD; ++b)
            if (b < nact) {
                // converged to a non-finite value -> failure (ssnode.py:257-262)
                a.status[(size_t)net * a.nb + b0 + b] = st[b];
                if (a.iters) a.iters[(size_t)net * a.nb + b0 + b] = its[b];
            }
    }
}

__global__ void ssn_status_fixup_f64_kernel(const double *R, int *status, int n_solves, int dim) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_solves) return;
    bool bad = false;
    for (int i = lane; i < dim; i += 32) bad |= !isfinite(R[(size_t)warp * dim + i]);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0 && bad && status[warp] == 0) status[warp] = 1;
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------

// opt-in shared memory per block of the CURRENT device (queried per call: cheap, and correct with several
// devices and threads, unlike a cached static)
static int max_optin_smem() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return v;
}
static bool force_smem_kernel() {
    const char *k1 = getenv("SSN_K1");
    return getenv("SSN_FORCE_SMEM_KERNEL") || (k1 && !strcmp(k1, "smem"));
}

typedef void (*FpKernel)(const FpArgs);
struct FpVariant { FpKernel fn; int threads; int rows; int kl; };

// rows covered = (32/KL) * NWARPS * TI; kpad must be a multiple of 4*KL
static const FpVariant kFpVariants[] = {
    {ssn_fp_cluster_kernel<4, 16, 8>, 256, 64, 16},
    {ssn_fp_cluster_kernel<7, 16, 8>, 256, 112, 16},
    {ssn_fp_cluster_kernel<7, 8, 8>, 256, 224, 8},
};

// Smallest cluster whose CTAs can hold their slice of the matrix, and the kernel
// instantiation that covers its rows.
bool choose_cluster_shape(int n_sites, ClusterShape *out, int smem_limit, int *variant) {
    const int dim = 2 * n_sites;
    ClusterShape s;
    s.dim = dim;
    for (int c = 1; c <= MAX_CLUSTER; c *= 2) {
        s.csize = c;
        s.rpc = (dim + c - 1) / c;
        if (s.rpc * (c - 1) >= dim) continue;               // a CTA would own no rows
        for (int v = 0; v < 3; ++v) {
            if (kFpVariants[v].rows < s.rpc) continue;
            const int q = 4 * kFpVariants[v].kl;
            s.kpad = ((dim + q - 1) / q) * q;
            if (smem_layout(s, n_sites).total <= smem_limit) { *out = s; *variant = v; return true; }
            break;
        }
    }
    return false;
}

struct FpLaunchPlan { FpVariant var; ClusterShape shape; int smem; int clusters; };

static int plan_fixed_point(int n_sites, int nz, FpLaunchPlan *plan) {
    const int limit = max_optin_smem();
    int variant = 0;
    if (!choose_cluster_shape(n_sites, &plan->shape, limit, &variant)) {
        set_error("fixed-point kernel: 2N=%d does not fit a cluster of %d CTAs (smem %d B)",
                  2 * n_sites, MAX_CLUSTER, limit);
        return -1;
    }
    plan->var = kFpVariants[variant];
    plan->smem = smem_layout(plan->shape, n_sites).total;
    SSN_CUDA(cudaFuncSetAttribute(plan->var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan->shape.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(plan->shape.csize, 1, 1);
    cfg.blockDim = dim3(plan->var.threads, 1, 1);
    cfg.dynamicSmemBytes = plan->smem;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    SSN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, plan->var.fn, &cfg));
    if (max_clusters < 1) {
        set_error("fixed-point kernel: no resident cluster of %d CTAs with %d B smem", plan->shape.csize, plan->smem);
        return -1;
    }
    plan->clusters = nz > 0 ? std::min(max_clusters, nz) : max_clusters;
    return 0;
}

int fixed_point_occupancy(int n_sites, int *cluster_size, int *resident_clusters) {
    if (!force_smem_kernel()) {
        ssn_solver sv = {};
        sv.io_type = SSN_IO_TANH; sv.k = 0.01; sv.n = 2.2; sv.rate_soft_bound = 200; sv.rate_hard_bound = 1000;
        if (ws_occupancy(sv, n_sites, cluster_size, resident_clusters) == 0) return 0;
    }
    FpLaunchPlan plan;
    int rc = plan_fixed_point(n_sites, 0, &plan);
    if (rc) return rc;
    if (cluster_size) *cluster_size = plan.shape.csize;
    if (resident_clusters) *resident_clusters = plan.clusters;
    return 0;
}

int fixed_point_kernel_name(int n_sites, char *buf, int cap) {
    if (!buf || cap < 1) return -1;
    if (!force_smem_kernel()) {
        ssn_solver sv = {};
        sv.io_type = SSN_IO_TANH; sv.k = 0.01; sv.n = 2.2; sv.rate_soft_bound = 200; sv.rate_hard_bound = 1000;
        if (ws_kernel_name(sv, n_sites, buf, cap) == 0) return 0;
    }
    FpLaunchPlan plan;
    int rc = plan_fixed_point(n_sites, 0, &plan);
    if (rc) return rc;
    snprintf(buf, cap, "ssn_fp_cluster_kernel<rows=%d,KL=%d>x%d", plan.var.rows, plan.var.kl, plan.shape.csize);
    return 0;
}

// All pointers are device pointers; `counter` is one int of scratch.
int launch_fixed_point_f32(const ssn_solver &sv, int nz, int nb, int n_sites, int w_kind, const float *w,
                           const ssn_jds *jds, const float *ext, int ext_per_network, const float *r_init,
                           float *R, int *status, int *iters, int *counter, cudaStream_t stream) {
    if (nz <= 0 || nb <= 0) return 0;
    const int n_solves_all = nz * nb;
    // Kernel choice: the warp-specialised register kernel, else (shapes out of its range return 1) the
    // shared-memory kernel.  SSN_K1=smem (or SSN_FORCE_SMEM_KERNEL) forces the latter.
    int rc = 1;
    if (!force_smem_kernel())
        rc = launch_fixed_point_ws(sv, nz, nb, n_sites, w_kind, w, jds, ext, ext_per_network, r_init, R, status, iters,
                                   counter, stream);
    if (rc == 0) {
        ssn_status_fixup_kernel<<<(n_solves_all * 32 + 255) / 256, 256, 0, stream>>>(R, status, n_solves_all, 2 * n_sites);
        SSN_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    if (rc != 1) return rc;
    FpLaunchPlan plan;
    rc = plan_fixed_point(n_sites, nz, &plan);
    if (rc) return rc;
    FpArgs a = {};
    a.nz = nz; a.nb = nb; a.n_sites = n_sites;
    a.shape = plan.shape;
    a.w_kind = w_kind; a.w = w;
    if (w_kind == SSN_W_FROM_Z) {
        if (!jds) { set_error("SSN_W_FROM_Z needs jds"); return -1; }
        a.wc = make_weight_const(*jds, n_sites);
    }
    a.ext = ext;
    a.ext_stride_z = ext_per_network ? (long long)nb * 2 * n_sites : 0;
    a.r_init = r_init;
    a.R = R; a.status = status; a.iters = iters; a.work_counter = counter;
    a.io = make_io_const<float>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.eps_E = sv.dt / sv.tau_E; a.eps_I = sv.dt / sv.tau_I;
    a.atol = sv.atol; a.r_hard = sv.rate_hard_bound;
    a.max_iter = sv.max_iter; a.check_hard = sv.io_type != SSN_IO_TANH;

    SSN_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.shape.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(plan.clusters * plan.shape.csize, 1, 1);
    cfg.blockDim = dim3(plan.var.threads, 1, 1);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {
        KernelTimer kt("ssn_fp_cluster_kernel", stream);
        SSN_CUDA(cudaLaunchKernelEx(&cfg, plan.var.fn, a));
    }
    count_launch();
    const int n_solves = nz * nb;
    ssn_status_fixup_kernel<<<(n_solves * 32 + 255) / 256, 256, 0, stream>>>(R, status, n_solves, 2 * n_sites);
    SSN_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_fixed_point_f64(const ssn_solver &sv, int nz, int nb, int n_sites, const double *W,
                           const double *ext, int ext_per_network, const double *r_init,
                           double *R, int *status, int *iters, bool nonfinite_fixup, int *counter, cudaStream_t stream) {
    if (nz <= 0 || nb <= 0) return 0;
    // W resident in cluster shared memory when it fits (SSN_F64=streamed forces the L2-streaming kernel)
    const char *f64 = getenv("SSN_F64");
    if (counter && !(f64 && !strcmp(f64, "streamed"))) {
        const int rc = launch_fixed_point_f64_cluster(sv, nz, nb, n_sites, W, ext, ext_per_network, r_init, R, status,
                                                      iters, counter, stream);
        if (rc == 0) {
            if (nonfinite_fixup) {
                const int n_solves = nz * nb;
                ssn_status_fixup_f64_kernel<<<(n_solves * 32 + 255) / 256, 256, 0, stream>>>(R, status, n_solves, 2 * n_sites);
                SSN_CUDA(cudaGetLastError());
                count_launch();
            }
            return 0;
        }
        if (rc != 1) return rc;
    }
    Fp64Args a = {};
    a.nz = nz; a.nb = nb; a.n_sites = n_sites; a.dim = 2 * n_sites;
    a.W = W; a.ext = ext;
    a.ext_stride_z = ext_per_network ? (long long)nb * 2 * n_sites : 0;
    a.r_init = r_init; a.R = R; a.status = status; a.iters = iters;
    a.io = make_io_const<double>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.eps_E = sv.dt / sv.tau_E; a.eps_I = sv.dt / sv.tau_I;
    a.atol = sv.atol; a.r_hard = sv.rate_hard_bound;
    a.max_iter = sv.max_iter; a.check_hard = sv.io_type != SSN_IO_TANH;
    const int tbd = nb == 1 ? 1 : 8;
    const int n_chunks = (nb + tbd - 1) / tbd;
    const size_t smem = (size_t)3 * a.dim * tbd * sizeof(double);
    if ((int)smem > max_optin_smem()) {
        set_error("float64 kernel: 2N=%d needs %zu B of shared memory", a.dim, smem);
        return -1;
    }
    {
        KernelTimer kt("ssn_fp64_kernel", stream);
        if (tbd == 1) {
            SSN_CUDA(cudaFuncSetAttribute(ssn_fp64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ssn_fp64_kernel<1><<<nz * n_chunks, 512, smem, stream>>>(a);
        } else {
            SSN_CUDA(cudaFuncSetAttribute(ssn_fp64_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ssn_fp64_kernel<8><<<nz * n_chunks, 512, smem, stream>>>(a);
        }
        SSN_CUDA(cudaGetLastError());
    }
    count_launch();
    const int n_solves = nz * nb;
    if (!nonfinite_fixup) return 0;
    ssn_status_fixup_f64_kernel<<<(n_solves * 32 + 255) / 256, 256, 0, stream>>>(R, status, n_solves, a.dim);
    SSN_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace ssn
