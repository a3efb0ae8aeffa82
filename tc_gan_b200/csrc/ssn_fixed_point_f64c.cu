// K1'': all-float64 fixed-point solve with the weight matrix resident in the shared memory of a cluster.
//
// This is the kernel behind the reference-ABI symbols solve_dynamics_asym_*_euler (one network x one
// stimulus per call, tc_gan/ext/ssnode.c:55-187) and behind `precise` batched calls.  A cluster of
// csize = ceil(2N / 56) CTAs (8 for 2N = 402) keeps W in DOUBLE in shared memory (402^2 * 8 B = 1.29 MB
// = 161 KB per CTA) for the life of a network; every CTA also keeps the whole state panel (2N x TBD
// doubles, double-buffered).  Per sweep: warp w contracts its 7 rows against the panel (FP64 FMA, lanes
// stride the columns), the 32 lanes reduce-scatter the 7 x TBD sums, the owner lanes apply the Euler step
// with the closed-form transfer function (pow / tanh in double, as the reference) and store the new state
// into the panel of every CTA of the cluster through distributed shared memory; one barrier.cluster per
// sweep.  Same stopping rule and codes as ssnode.c:84-102.
//
// The older ssn_fp64_kernel (one CTA per panel, W streamed from L2 every sweep: 40 us per sweep) remains
// for shapes whose slice does not fit in shared memory.
#include <algorithm>
#include <cstdlib>
#include "ssn_cluster_core.cuh"
#include "ssn_launch.h"

namespace ssn {

constexpr int FC_WARPS = 8, FC_THREADS = 32 * FC_WARPS, FC_TI = 7;
#ifndef FC_UNROLL_1
#define FC_UNROLL_1 4
#endif
#ifndef FC_UNROLL_8
#define FC_UNROLL_8 1
#endif

struct Fc64Args {
    int nz, nb, n_sites, dim, csize, rpc;
    const double *W;               // [nz][dim][dim]
    const double *ext;             // [nb][dim] or [nz][nb][dim]
    long long ext_stride_z;
    const double *r_init;          // [nz][nb][dim] or null
    double *R;                     // [nz][nb][dim]
    int *status, *iters;
    int *work_counter;
    IoConst<double> io;
    double eps_E, eps_I, atol, r_hard;
    int max_iter, check_hard;
};

struct Fc64Misc {
    unsigned flags[2][MAX_CLUSTER];
    unsigned myflags[2];
    int next_item;
};

__host__ __device__ inline int fc_x_doubles(int dim, int tbd) { return 2 * dim * tbd; }      // [2 buffers][planes][dim][1 or 2]
__host__ __device__ inline size_t fc_smem_bytes(int dim, int rpc, int tbd) {
    return ((size_t)rpc * dim + fc_x_doubles(dim, tbd)) * 8 + 256;
}

__device__ __forceinline__ void st_cluster_f64(unsigned addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Transfer function in double (tc_gan/ext/ssnode.c:25-53).  k v^n is evaluated as k exp(n log v): within 2e-15
// (relative) of pow() over the whole range and about half its instruction count.
__device__ __forceinline__ double io_eval_f64(const IoConst<double> &io, double v) {
    if (!(v > 0.0)) return v != v ? v : 0.0;
    if (io.io_type != SSN_IO_POWER && v > io.v0)
        return io.io_type == SSN_IO_LINEAR ? fma(io.lin_slope, v - io.v0, io.r_soft)
                                           : io.r_soft + io.span * tanh(io.tanh_scale * (v - io.v0));
    return io.k * exp(io.n * log(v));
}

// TBD = 1: panel X[buf][j];  TBD = 8: four planes of stimulus pairs, X[buf][pair][j] as double2.
template <int TBD>
__global__ void __launch_bounds__(FC_THREADS, 1) ssn_fp64_cluster_kernel(const Fc64Args a) {
    static_assert(TBD == 1 || TBD == 8, "panel width");
    constexpr int NOWN = TBD == 8 ? 2 : 1;                       // outputs per owner lane
    constexpr int FC_UNROLL = TBD == 8 ? FC_UNROLL_8 : FC_UNROLL_1;   // column steps whose loads are issued together
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dim = a.dim, csize = a.csize, rpc = a.rpc, N = a.n_sites;
    double *Wsm = reinterpret_cast<double *>(smem);                               // [rpc][dim]
    double *X = Wsm + (size_t)rpc * dim;                                          // panel buffers
    Fc64Misc *misc = reinterpret_cast<Fc64Misc *>(X + fc_x_doubles(dim, TBD));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = rank * rpc;
    const int rows_here = max(0, min(rpc, dim - row_base));
    const int row0 = warp * FC_TI;

    // ownership after the reduction: TBD = 8 as in the FP32 kernels (lane bit 4 -> stimulus half, bits 3..1 ->
    // row, bit 0 -> stimulus pair of the half); TBD = 1: lane t owns row t
    const int my_t = TBD == 8 ? (lane >> 1) & 7 : lane;
    const int my_st0 = TBD == 8 ? 4 * (lane >> 4) + 2 * (lane & 1) : 0;
    const bool owner = my_t < FC_TI && row0 + my_t < rows_here;
    const int grow = row_base + row0 + my_t;
    const double eps_own = grow < N ? a.eps_E : a.eps_I;
    // element (buffer, column j, stimulus b) of the panel, in doubles
    auto xidx = [&](int buf, int j, int b) { return TBD == 8 ? ((buf * 4 + (b >> 1)) * dim + j) * 2 + (b & 1) : buf * dim + j; };
    const unsigned x_local = smem_u32(X);
    unsigned pdelta[MAX_CLUSTER];
#pragma unroll
    for (int p = 0; p < MAX_CLUSTER; ++p) pdelta[p] = map_to_rank(x_local, (unsigned)(p < csize ? p : 0)) - x_local;
    if (tid < 2) misc->myflags[tid] = 0u;
    if (tid < 2 * MAX_CLUSTER) misc->flags[tid / MAX_CLUSTER][tid % MAX_CLUSTER] = 0u;
    cluster.sync();

    const int n_chunks = (a.nb + TBD - 1) / TBD;
    const int total = a.nz * n_chunks;
    const unsigned all = TBD == 8 ? 0xffu : 1u;
    int cur_net = -1;

    for (;;) {
        if (rank == 0 && tid == 0) {
            const int n = atomicAdd(a.work_counter, 1);
            for (int p = 0; p < csize; ++p) st_cluster_u32(map_to_rank(smem_u32(&misc->next_item), p), (unsigned)n);
        }
        cluster.sync();
        const int item = misc->next_item;
        if (item >= total) break;
        const int net = item / n_chunks, chunk = item - net * n_chunks;
        const int b0 = chunk * TBD, nact = min(TBD, a.nb - b0);

        if (net != cur_net) {                                   // this CTA's rows of W -> shared memory
            const double *src = a.W + (size_t)net * dim * dim + (size_t)row_base * dim;
            for (int i = tid; i < rows_here * dim; i += FC_THREADS) Wsm[i] = src[i];
            cur_net = net;
        }
        // every CTA fills its own copy of the initial panel
        for (int i = tid; i < dim * TBD; i += FC_THREADS) {
            const int j = i / TBD, b = i - j * TBD;
            X[xidx(0, j, b)] = (b < nact && a.r_init) ? a.r_init[((size_t)net * a.nb + b0 + b) * dim + j] : 0.0;
        }
        double r_cur[NOWN], e_own[NOWN];
#pragma unroll
        for (int q = 0; q < NOWN; ++q) {
            const int b = my_st0 + q;
            const bool ok = owner && b < nact;
            e_own[q] = ok ? a.ext[(size_t)net * a.ext_stride_z + (size_t)(b0 + b) * dim + grow] : 0.0;
            r_cur[q] = (ok && a.r_init) ? a.r_init[((size_t)net * a.nb + b0 + b) * dim + grow] : 0.0;
        }
        __syncthreads();

        unsigned done = nact >= TBD ? 0u : (all & ~((1u << nact) - 1u));
        int st[TBD], its[TBD];
#pragma unroll
        for (int b = 0; b < TBD; ++b) { st[b] = 1; its[b] = a.max_iter; }
        int buf = 0;

        for (int it = 1; it <= a.max_iter; ++it) {
            // ---- contraction: 7 rows x TBD stimuli per warp, lanes stride the columns ----
            double acc[FC_TI][TBD];
#pragma unroll
            for (int t = 0; t < FC_TI; ++t)
#pragma unroll
                for (int b = 0; b < TBD; ++b) acc[t][b] = 0.0;
            const double *wrow = Wsm + (size_t)min(row0, max(rows_here - 1, 0)) * dim;
#pragma unroll 1
            for (int j0 = 0; j0 < dim; j0 += 32 * FC_UNROLL)
#pragma unroll
            for (int jj = 0; jj < FC_UNROLL; ++jj) {
                const int j = j0 + 32 * jj + lane;
                if (j >= dim) continue;
                double xv[TBD];
                if (TBD == 8) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const double2 v = *reinterpret_cast<const double2 *>(X + xidx(buf, j, 2 * p));
                        xv[2 * p] = v.x; xv[2 * p + 1] = v.y;
                    }
                } else {
                    xv[0] = X[xidx(buf, j, 0)];
                }
#pragma unroll
                for (int t = 0; t < FC_TI; ++t) {
                    const double w = row0 + t < rows_here ? wrow[(size_t)t * dim + j] : 0.0;
#pragma unroll
                    for (int b = 0; b < TBD; ++b) acc[t][b] = fma(w, xv[b], acc[t][b]);
                }
            }
            // ---- reduction over the 32 lanes ----
            double dv[NOWN];
            if (TBD == 8) {
                const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
                double h8[8][4];
#pragma unroll
                for (int t = 0; t < FC_TI; ++t)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double send = u16 ? acc[t][c] : acc[t][4 + c];
                        const double keep = u16 ? acc[t][4 + c] : acc[t][c];
                        h8[t][c] = keep + shfl_xor_f64(send, 16);
                    }
#pragma unroll
                for (int c = 0; c < 4; ++c) h8[7][c] = 0.0;
                double h4[4][4], h2[2][4], h1[4];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double send = u8 ? h8[t][c] : h8[4 + t][c];
                        const double keep = u8 ? h8[4 + t][c] : h8[t][c];
                        h4[t][c] = keep + shfl_xor_f64(send, 8);
                    }
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double send = u4 ? h4[t][c] : h4[2 + t][c];
                        const double keep = u4 ? h4[2 + t][c] : h4[t][c];
                        h2[t][c] = keep + shfl_xor_f64(send, 4);
                    }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double send = u2 ? h2[0][c] : h2[1][c];
                    const double keep = u2 ? h2[1][c] : h2[0][c];
                    h1[c] = keep + shfl_xor_f64(send, 2);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double send = u1 ? h1[c] : h1[2 + c];
                    const double keep = u1 ? h1[2 + c] : h1[c];
                    dv[c] = keep + shfl_xor_f64(send, 1);
                }
            } else {
                double mine = 0.0;
#pragma unroll
                for (int t = 0; t < FC_TI; ++t) {
                    double v = acc[t][0];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_f64(v, o);
                    mine = t == lane ? v : mine;
                }
                dv[0] = mine;
            }
            // ---- Euler step of the owned outputs, published to every CTA of the cluster ----
            unsigned word = 0u;
            const int nbuf = buf ^ 1;
            if (owner) {
#pragma unroll
                for (int q = 0; q < NOWN; ++q) {
                    const int b = my_st0 + q;
                    if (b < nact) {
                        double r_new = r_cur[q];
                        if (!((done >> b) & 1u)) {
                            const double fv = io_eval_f64(a.io, dv[q] + e_own[q]);
                            r_new = r_cur[q] + (fv - r_cur[q]) * eps_own;
                            if (fabs(r_new - r_cur[q]) >= a.atol) word |= 1u << b;
                            if (r_new >= a.r_hard) word |= 1u << (8 + b);
                            r_cur[q] = r_new;
                        }
                        const unsigned off = 8u * (unsigned)xidx(nbuf, grow, b);
#pragma unroll
                        for (int p = 0; p < MAX_CLUSTER; ++p)
                            if (p < csize) st_cluster_f64(x_local + off + pdelta[p], r_new);
                    }
                }
            }
            word = __reduce_or_sync(0xffffffffu, word);
            if (lane == 0 && word) atomicOr(&misc->myflags[it & 1], word);
            __syncthreads();
            if (tid < csize)
                st_cluster_u32(map_to_rank(smem_u32(&misc->flags[it & 1][rank]), (unsigned)tid), misc->myflags[it & 1]);
            if (tid == 32) misc->myflags[(it + 1) & 1] = 0u;
            cluster.sync();
            unsigned F = 0u;
#pragma unroll
            for (int p = 0; p < MAX_CLUSTER; ++p) F |= p < csize ? misc->flags[it & 1][p] : 0u;
            const unsigned conv_now = ~(F & 0xffu) & ~done & all;                          // ssnode.c:84-96 first ...
            const unsigned hard_now = a.check_hard ? ((F >> 8) & ~done & ~conv_now & all) : 0u;   // ... then :98-102
#pragma unroll
            for (int b = 0; b < TBD; ++b) {
                if ((conv_now >> b) & 1u) { st[b] = 0; its[b] = it; }
                if ((hard_now >> b) & 1u) { st[b] = 2; its[b] = it; }
            }
            done |= conv_now | hard_now;
            buf = nbuf;
            if (done == all) break;
        }
        // ---- results ----
        if (owner) {
#pragma unroll
            for (int q = 0; q < NOWN; ++q) {
                const int b = my_st0 + q;
                if (b < nact) a.R[((size_t)net * a.nb + b0 + b) * dim + grow] = r_cur[q];
            }
        }
        if (rank == 0 && tid == 0) {
#pragma unroll
            for (int b = 0; b < TBD; ++b)
                if (b < nact) {
                    a.status[(size_t)net * a.nb + b0 + b] = st[b];
                    if (a.iters) a.iters[(size_t)net * a.nb + b0 + b] = its[b];
                }
        }
    }
}

// Returns 1 when the shape does not fit (the caller then streams W from L2), 0 on success.
int launch_fixed_point_f64_cluster(const ssn_solver &sv, int nz, int nb, int n_sites, const double *W,
                                   const double *ext, int ext_per_network, const double *r_init,
                                   double *R, int *status, int *iters, int *counter, cudaStream_t stream) {
    const int dim = 2 * n_sites, rows = FC_TI * FC_WARPS;
    const int csize = (dim + rows - 1) / rows;
    if (csize > MAX_CLUSTER) return 1;
    const int rpc = (dim + csize - 1) / csize;
    if (rpc * (csize - 1) >= dim) return 1;
    const int tbd = nb == 1 ? 1 : 8;
    const size_t smem = fc_smem_bytes(dim, rpc, tbd);
    int dev = 0, limit = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)limit) return 1;
    void (*fn)(const Fc64Args) = tbd == 1 ? ssn_fp64_cluster_kernel<1> : ssn_fp64_cluster_kernel<8>;
    SSN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(csize, 1, 1);
    cfg.blockDim = dim3(FC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg) != cudaSuccess) { cudaGetLastError(); return 1; }
    if (max_clusters < 1) return 1;
    const int n_chunks = (nb + tbd - 1) / tbd;
    const int clusters = std::min(max_clusters, nz * n_chunks);
    cfg.gridDim = dim3(clusters * csize, 1, 1);

    Fc64Args a = {};
    a.nz = nz; a.nb = nb; a.n_sites = n_sites; a.dim = dim; a.csize = csize; a.rpc = rpc;
    a.W = W; a.ext = ext; a.ext_stride_z = ext_per_network ? (long long)nb * dim : 0;
    a.r_init = r_init; a.R = R; a.status = status; a.iters = iters; a.work_counter = counter;
    a.io = make_io_const<double>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.eps_E = sv.dt / sv.tau_E; a.eps_I = sv.dt / sv.tau_I;
    a.atol = sv.atol; a.r_hard = sv.rate_hard_bound;
    a.max_iter = sv.max_iter; a.check_hard = sv.io_type != SSN_IO_TANH;
    SSN_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    {
        KernelTimer kt("ssn_fp64_cluster_kernel", stream);
        SSN_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
    }
    count_launch();
    return 0;
}

}  // namespace ssn
