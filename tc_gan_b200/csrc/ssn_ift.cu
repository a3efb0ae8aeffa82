// K2: implicit-function-theorem generator gradient at the SSN fixed points.
//
// Replaces tc_gan/gradient_expressions/SS_grad.py:17-76 (WRgrad_batch: one dense
// MatrixInverse per (network, stimulus)), make_w_batch.py:36-121 (the
// [nz,2N,2N,2,2] dW/dtheta tensors) and the contraction of run/gan.py:902-911 by
// the adjoint form
//     (I - W^T Phi) mu = g,   dL/dW = (Phi mu) r^T,   dL/dtheta = <dL/dW, dW/dtheta>
// with Phi = diag f'(W r + I) and g = dL/dr.  The adjoint system is solved by the
// damped iteration  mu <- mu + eps (g - mu + W^T (Phi mu)),  eps = dt/tau, whose
// iteration matrix has the spectrum of the forward Euler linearisation, on the
// same cluster-resident machinery as K1 (here the cluster holds W^T).
//
// Per network and 8-stimulus panel:  (1) W in smem, one contraction v = W r + I
// -> Phi;  (2) W^T in smem, adjoint sweeps until max|d mu| < rtol * max|g|;
// (3) fused reduction of (Phi mu)_i r_j against dW_ij/dtheta (z re-read from
// global, never materialising dL/dW) into 12 doubles.
#include "ssn_cluster_core.cuh"
#include "ssn_launch.h"

namespace ssn {

struct IftArgs {
    int nz, nb, n_sites;
    ClusterShape shape;
    const float *z;
    WeightConst wc;
    const float *ext;
    long long ext_stride_z;
    const float *R, *g;
    float *mu, *grad_ext;
    int *status, *iters;
    double *grad;                  // [12] J, D, S
    int *work_counter;
    IoConst<float> io;
    double eps_E, eps_I, rtol;
    int max_iter;
};

template <int TI, int KL, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) ssn_ift_cluster_kernel(const IftArgs a) {
    using Own = Owner<TI, KL>;
    constexpr int TO = Own::TO;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[12];
    __shared__ unsigned gmax_bits[TB];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int csize = a.shape.csize;
    const int dim = a.shape.dim, kpad = a.shape.kpad, rpc = a.shape.rpc, N = a.n_sites;
    const int P = panel_P(kpad);
    const SmemLayout L = smem_layout(a.shape, a.n_sites);
    float *Wsm = reinterpret_cast<float *>(smem + L.w_off);
    float *Xf = reinterpret_cast<float *>(smem + L.x_off);
    const float4 *X4 = reinterpret_cast<const float4 *>(smem + L.x_off);
    float *rsm = reinterpret_cast<float *>(smem + L.ext_off);          // [local row][8] staging
    float *gtab = reinterpret_cast<float *>(smem + L.gtab_off);
    Misc *misc = reinterpret_cast<Misc *>(smem + L.misc_off);

    const int tid = threadIdx.x, nthreads = NWARPS * 32;
    const int warp = tid >> 5, lane = tid & 31;
    const int kl = lane % KL;
    const int grp = warp * (32 / KL) + lane / KL;
    const int row_base = rank * rpc;
    const int rows_here = max(0, min(rpc, dim - row_base));

    int wrow[TI];
#pragma unroll
    for (int t = 0; t < TI; ++t) wrow[t] = min(grp * TI + t, rows_here - 1);
    const int my_stim = Own::stim(kl);
    const int own0 = grp * TI + Own::first_row(kl);
    bool valid[TO];
#pragma unroll
    for (int u = 0; u < TO; ++u)
        valid[u] = (Own::first_row(kl) + u < TI) && (own0 + u < rows_here);

    unsigned xpeer[MAX_CLUSTER];
#pragma unroll
    for (int p = 0; p < MAX_CLUSTER; ++p) xpeer[p] = map_to_rank(smem_u32(Xf), p < csize ? p : 0);

    build_profile_table(a.wc, N, gtab, tid, nthreads);
    for (int i = tid; i < 2 * 2 * P * 4; i += nthreads) Xf[i] = 0.f;
    if (tid == 0) misc->myflags = 0u;
    __syncthreads();

    const int n_chunks = (a.nb + TB - 1) / TB;
    const unsigned buf_bytes = 2u * (unsigned)P * 16u;

    for (;;) {
        if (rank == 0 && tid == 0) {
            const int n = atomicAdd(a.work_counter, 1);
            for (int p = 0; p < csize; ++p) st_cluster_u32(map_to_rank(smem_u32(&misc->next_net), p), (unsigned)n);
        }
        cluster.sync();
        const int net = misc->next_net;
        if (net >= a.nz) break;
        const float *z_net = a.z + (size_t)net * dim * dim;
        const float *ext_net = a.ext + (size_t)net * a.ext_stride_z;

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int b0 = chunk * TB;
            const int nact = min(TB, a.nb - b0);
            const bool active = my_stim < nact;
            const size_t sol = (size_t)net * a.nb + b0 + my_stim;       // this thread's solve

            // ---------- phase 1: Phi = f'(W r + I) ----------
            load_matrix_slice<false>(Wsm, z_net, SSN_W_FROM_Z, a.wc, gtab, N, dim, kpad, row_base, rows_here,
                                     tid, nthreads);
            if (tid < TB) gmax_bits[tid] = 0u;
            if (tid < 12) red[tid] = 0.0;
            unsigned xoff[TO];
            float phi[TO], g_own[TO];
            double eps_own[TO], mu[TO];
#pragma unroll
            for (int u = 0; u < TO; ++u) {
                xoff[u] = 0u; phi[u] = 0.f; g_own[u] = 0.f; eps_own[u] = 0.0; mu[u] = 0.0;
                if (valid[u]) {
                    const int gr = row_base + own0 + u;
                    xoff[u] = 4u * (unsigned)panel_index(P, 0, gr, my_stim);
                    eps_own[u] = gr < N ? a.eps_E : a.eps_I;
                    const float rf = active ? __ldg(a.R + sol * dim + gr) : 0.f;
#pragma unroll
                    for (int p = 0; p < MAX_CLUSTER; ++p)
                        if (p < csize) st_cluster_f32(xpeer[p] + xoff[u], rf);
                }
            }
            cluster.sync();
            {
                float acc[TI][TB], v[TO];
                contract_panel<TI, KL>(acc, Wsm, X4, P, kpad, 0, wrow, kl);
                reduce_scatter<TI, KL>(acc, v, kl);
                float gm = 0.f;
#pragma unroll
                for (int u = 0; u < TO; ++u)
                    if (valid[u] && active) {
                        const int gr = row_base + own0 + u;
                        const float e = __ldg(ext_net + (size_t)(b0 + my_stim) * dim + gr);
                        phi[u] = io_gain<float>(a.io, v[u] + e);
                        if (!isfinite(phi[u])) phi[u] = 0.f;          // non-finite state of a rejected solve
                        g_own[u] = __ldg(a.g + sol * dim + gr);
                        mu[u] = (double)g_own[u];
                        gm = fmaxf(gm, fabsf(g_own[u]));
                    }
                if (active && gm > 0.f) atomicMax(&gmax_bits[my_stim], __float_as_uint(gm));
            }
            __syncthreads();                       // everyone is done reading W from smem
            if (tid == 0)
                for (int b = 0; b < TB; ++b)
                    for (int p = 0; p < csize; ++p)
                        st_cluster_u32(map_to_rank(smem_u32(&misc->scale[0][rank][b]), p), gmax_bits[b]);

            // ---------- phase 2: adjoint sweeps with W^T ----------
            load_matrix_slice<true>(Wsm, z_net, SSN_W_FROM_Z, a.wc, gtab, N, dim, kpad, row_base, rows_here,
                                    tid, nthreads);
#pragma unroll
            for (int u = 0; u < TO; ++u)
                if (valid[u]) {
                    const float af = (float)(phi[u] * mu[u]);
#pragma unroll
                    for (int p = 0; p < MAX_CLUSTER; ++p)
                        if (p < csize) st_cluster_f32(xpeer[p] + xoff[u] + buf_bytes, af);   // buffer 1
                }
            cluster.sync();
            float gscale = 0.f;                    // max|g| of my stimulus over the whole network
            for (int p = 0; p < csize; ++p)
                gscale = fmaxf(gscale, __uint_as_float(reinterpret_cast<const unsigned *>(&misc->scale[0][p][0])[my_stim]));
            const double tol = a.rtol * (double)fmaxf(gscale, 1e-30f);

            unsigned done = nact >= TB ? 0u : (0xffu << nact) & 0xffu;
            int my_status = 1, my_iters = a.max_iter;
            // A solve whose dL/dr vanishes (e.g. a rejected network masked out by the caller) has mu = 0: it is
            // finished before the first sweep and contributes nothing, whatever (possibly non-finite) state R holds.
            unsigned dead = 0u;
            for (int b = 0; b < TB; ++b) {
                float gs = 0.f;
                for (int p = 0; p < csize; ++p) gs = fmaxf(gs, __uint_as_float(reinterpret_cast<const unsigned *>(&misc->scale[0][p][0])[b]));
                if (!(gs > 0.f)) dead |= 1u << b;
            }
            done |= dead;
            if ((dead >> my_stim) & 1u) {
                my_status = 0; my_iters = 0;
#pragma unroll
                for (int u = 0; u < TO; ++u) { mu[u] = 0.0; phi[u] = 0.f; g_own[u] = 0.f; }
            }
            int buf = 1;
            for (int it = 1; it <= a.max_iter && done != 0xffu; ++it) {
                float acc[TI][TB], y[TO];
                contract_panel<TI, KL>(acc, Wsm, X4, P, kpad, buf, wrow, kl);
                reduce_scatter<TI, KL>(acc, y, kl);
                const int nbuf = buf ^ 1;
                const bool frozen = (done >> my_stim) & 1u;
                bool moving = false;
#pragma unroll
                for (int u = 0; u < TO; ++u)
                    if (valid[u]) {
                        const double m_old = mu[u];
                        const double m_new = m_old + ((double)g_own[u] - m_old + (double)y[u]) * eps_own[u];
                        if (!frozen) {
                            moving |= fabs(m_new - m_old) >= tol;
                            mu[u] = m_new;
                        }
                        const float af = (float)((double)phi[u] * mu[u]);
                        const unsigned off = xoff[u] + (nbuf ? buf_bytes : 0u);
#pragma unroll
                        for (int p = 0; p < MAX_CLUSTER; ++p)
                            if (p < csize) st_cluster_f32(xpeer[p] + off, af);
                    }
                const unsigned mm = stim_mask<KL>(moving);
                if (lane == 0 && mm) atomicOr(&misc->myflags, mm);
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
                const unsigned conv_now = ~F & ~done & 0xffu;
                if ((conv_now >> my_stim) & 1u) { my_status = 0; my_iters = it; }
                done |= conv_now;
                buf = nbuf;
                if (done == 0xffu) break;
            }

            // ---------- results of the solve ----------
#pragma unroll
            for (int u = 0; u < TO; ++u)
                if (valid[u] && active) {
                    if (a.mu) a.mu[sol * dim + row_base + own0 + u] = (float)mu[u];
                    if (a.grad_ext) a.grad_ext[sol * dim + row_base + own0 + u] = (float)((double)phi[u] * mu[u]);
                }
            if (rank == 0 && tid < KL && (kl % Own::SPLIT) == 0 && active) {
                if (a.status) a.status[sol] = my_status;
                if (a.iters) a.iters[sol] = my_iters;
            }

            // ---------- phase 3: dL/dtheta += sum_b <(Phi mu)_b r_b^T, dW/dtheta> over my columns of W ----------
            // This CTA owns rows of W^T, i.e. columns j of W.  Panel `buf` holds (Phi mu)_i for every i.
            for (int i = tid; i < rows_here * TB; i += nthreads) {
                const int r = i / TB, b = i % TB;
                rsm[i] = (b < nact && !((dead >> b) & 1u)) ? __ldg(a.R + ((size_t)net * a.nb + b0 + b) * dim + row_base + r) : 0.f;
            }
            __syncthreads();
            const float4 *A0 = X4 + (buf * 2 + 0) * P, *A1 = A0 + P;
            for (int ah = 0; ah < 2; ++ah)                  // row half of W (index i)
                for (int bh = 0; bh < 2; ++bh) {            // column half of W (index j, mine)
                    const int j_lo = max(row_base, bh * N), j_hi = min(row_base + rows_here, (bh + 1) * N);
                    const int nj = j_hi - j_lo;
                    float sJ = 0.f, sD = 0.f, sS = 0.f;
                    if (nj > 0) {
                        const int ab = ah * 2 + bh;
                        const float cJ = a.wc.sJ[ab], cD = a.wc.sD[ab];
                        const float *gt = gtab + ab * N;
#pragma unroll 2
                        for (int idx = tid; idx < nj * N; idx += nthreads) {
                            const int ii = idx / nj, jl = idx - ii * nj;
                            const int i = ah * N + ii, j = j_lo + jl;
                            const float zz = __ldg(z_net + (size_t)i * dim + j);
                            int d = ii - (j - bh * N);
                            d = d < 0 ? -d : d;
                            const float gq = gt[d];
                            const float4 a0 = A0[panel_col(i)], a1 = A1[panel_col(i)];
                            const float4 r0 = *reinterpret_cast<const float4 *>(rsm + (j - row_base) * TB);
                            const float4 r1 = *reinterpret_cast<const float4 *>(rsm + (j - row_base) * TB + 4);
                            const float G = a0.x * r0.x + a0.y * r0.y + a0.z * r0.z + a0.w * r0.w +
                                            a1.x * r1.x + a1.y * r1.y + a1.z * r1.z + a1.w * r1.w;
                            const float x = (float)d * a.wc.dx;
                            const float gG = gq * G;
                            sJ += gG;
                            sD = fmaf(gG, zz, sD);
                            sS = fmaf(gG * x * x, fmaf(cD, zz, cJ), sS);
                        }
                        const float sgn = bh == 0 ? 1.f : -1.f;
                        sJ *= sgn; sD *= sgn; sS *= a.wc.invS3[ab];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sJ += __shfl_xor_sync(0xffffffffu, sJ, o);
                        sD += __shfl_xor_sync(0xffffffffu, sD, o);
                        sS += __shfl_xor_sync(0xffffffffu, sS, o);
                    }
                    if (lane == 0 && nj > 0) {
                        const int ab = ah * 2 + bh;
                        atomicAdd(&red[ab], (double)sJ);
                        atomicAdd(&red[4 + ab], (double)sD);
                        atomicAdd(&red[8 + ab], (double)sS);
                    }
                }
            __syncthreads();
            if (tid < 12) atomicAdd(a.grad + tid, red[tid]);
            // panels, rsm and Wsm are reused by the next panel / network
            cluster.sync();
        }
    }
}

typedef void (*IftKernel)(const IftArgs);
struct IftVariant { IftKernel fn; int threads; int rows; int kl; };
static const IftVariant kIftVariants[] = {
    {ssn_ift_cluster_kernel<4, 16, 8>, 256, 64, 16},
    {ssn_ift_cluster_kernel<7, 16, 8>, 256, 112, 16},
    {ssn_ift_cluster_kernel<7, 8, 8>, 256, 224, 8},
};

bool choose_cluster_shape(int n_sites, ClusterShape *out, int smem_limit, int *variant);

int launch_ift_gradient(const ssn_solver &sv, int nz, int nb, int n_sites, const float *z, const ssn_jds &jds,
                        const float *ext, int ext_per_network, const float *R, const float *g, double rtol,
                        double *grad, float *mu, int *status, int *iters, float *grad_ext, int *counter,
                        cudaStream_t stream) {
    SSN_CUDA(cudaMemsetAsync(grad, 0, 12 * sizeof(double), stream));
    if (nz <= 0 || nb <= 0) return 0;
    int dev = 0, limit = 0, variant = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    IftArgs a = {};
    if (!choose_cluster_shape(n_sites, &a.shape, limit - 256, &variant)) {
        set_error("ift kernel: 2N=%d does not fit a cluster of %d CTAs", 2 * n_sites, MAX_CLUSTER);
        return -1;
    }
    const IftVariant var = kIftVariants[variant];
    const int smem = smem_layout(a.shape, n_sites).total;
    SSN_CUDA(cudaFuncSetAttribute(var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    a.nz = nz; a.nb = nb; a.n_sites = n_sites;
    a.z = z; a.wc = make_weight_const(jds, n_sites);
    a.ext = ext; a.ext_stride_z = ext_per_network ? (long long)nb * 2 * n_sites : 0;
    a.R = R; a.g = g; a.mu = mu; a.grad_ext = grad_ext; a.status = status; a.iters = iters; a.grad = grad; a.work_counter = counter;
    a.io = make_io_const<float>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.eps_E = sv.dt / sv.tau_E; a.eps_I = sv.dt / sv.tau_I;
    a.rtol = rtol > 0 ? rtol : 1e-6;
    a.max_iter = sv.max_iter;

    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.shape.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(a.shape.csize, 1, 1);
    cfg.blockDim = dim3(var.threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    SSN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, var.fn, &cfg));
    if (max_clusters < 1) { set_error("ift kernel: no resident cluster"); return -1; }
    cfg.gridDim = dim3(std::min(max_clusters, nz) * a.shape.csize, 1, 1);
    SSN_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    {
        KernelTimer kt("ssn_ift_cluster_kernel", stream);
        SSN_CUDA(cudaLaunchKernelEx(&cfg, var.fn, a));
    }
    count_launch();
    return 0;
}

}  // namespace ssn
