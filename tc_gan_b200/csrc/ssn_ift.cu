// K2: implicit-function-theorem generator gradient at the SSN fixed points.
//
// Replaces tc_gan/gradient_expressions/SS_grad.py:17-76 (WRgrad_batch: one dense
// MatrixInverse per (network, stimulus)), make_w_batch.py:36-121 (the
// [nz,2N,2N,2,2] dW/dtheta tensors) and the contraction of run/gan.py:902-911 by
// the adjoint form
//     (I - W^T Phi) mu = g,   dL/dW = (Phi mu) r^T,   dL/dtheta = <dL/dW, dW/dtheta>
// with Phi = diag f'(W r + I) and g = dL/dr.  The adjoint system is solved by restarted GMRES(16) on the
// cluster-resident machinery of K1 (here the cluster holds W^T in shared memory): one Arnoldi step is one
// skinny contraction W^T (Phi v) of the 8-stimulus panel -- the eight systems of a network share W^T, differ in
// Phi and run in lockstep -- plus two cluster-wide reductions (Gram-Schmidt coefficients; norm of the new vector, whose
// barrier also carries the publish of the next direction).  The
// Krylov basis lives in an L2-resident scratch (16 vectors x 16 KB per resident cluster), the small least-squares problem
// is updated with Givens rotations per stimulus, and every cycle starts from the TRUE residual g - A mu, which
// is also the stopping test: ||g - A mu||_2 <= rtol ||g||_2.  A solve on which two consecutive cycles make no
// progress above 16 rtol (restarted GMRES can stall) continues with damped steps from the iterate it has.  Against the damped adjoint iteration
// mu <- mu + eps (g - mu + W^T Phi mu) of round 1 (still here: SSN_IFT=damped) this needs ~45 instead of ~500
// sweeps per panel at 2N = 402 (27 instead of 417 per solve) and ends 20x closer to the exact solution (2e-6
// instead of 4e-5).
//
// Per network and 8-stimulus panel:  (1) W in smem, one contraction v = W r + I -> Phi;  (2) W^T in smem,
// GMRES;  (3) fused reduction of (Phi mu)_i r_j against dW_ij/dtheta (z re-read from global, never
// materialising dL/dW) into 12 doubles.
#include <atomic>
#include <cstdlib>
#include <cstring>
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
    float4 *basis;                 // GMRES: [clusters][GM][TO4][csize][threads] Krylov vectors (L2-resident scratch)
    int extra_off;                 // GMRES: byte offset of the reduction / least-squares scratch in shared memory
    int test_stall;                // development (SSN_IFT=stall): treat every unfinished cycle as a stall
};

constexpr int GM = 16;                                 // GMRES restart length
constexpr int GNV = GM + 1;                            // values per stimulus in one cluster reduction: h_0..h_j, |w|^2
constexpr int RPK = GM * (GM + 1) / 2;                 // packed upper-triangular R of the rotated Hessenberg matrix
constexpr int IFT_NWARPS = 8;
// shared-memory scratch of the GMRES path, in floats: per-warp partial sums, per-CTA partial sums of every peer, the
// reduced values, R and the rotated right-hand side per stimulus, two control words
constexpr int IFT_EXTRA_FLOATS = IFT_NWARPS * TB * GNV + MAX_CLUSTER * TB * GNV + TB * GNV + TB * RPK + TB * (GM + 1) + 4 + TB * GNV;
constexpr int IFT_EXTRA_SMEM = ((IFT_EXTRA_FLOATS * 4 + 15) / 16) * 16;

// sum over the lanes of a warp that own the same stimulus (lane l owns stimulus (l % KL) / (KL / 8))
template <int KL>
__device__ __forceinline__ float stim_lane_sum(float p) {
#pragma unroll
    for (int o = 1; o < KL / TB; o <<= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
#pragma unroll
    for (int o = KL; o < 32; o <<= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    return p;
}

template <int TI, int KL, int NWARPS, bool GMRES>
__global__ void __launch_bounds__(NWARPS * 32, 1) ssn_ift_cluster_kernel(const IftArgs a) {
    using Own = Owner<TI, KL>;
    constexpr int TO = Own::TO;
    constexpr int TO4 = (TO + 3) / 4;
    static_assert(NWARPS == IFT_NWARPS, "scratch layout");
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
            int my_status = 1, my_iters = 0, buf = 0;
            unsigned dead = 0u;
            if constexpr (GMRES) {
                // One panel buffer is enough here (a publish always follows a cluster barrier that every CTA reaches
                // after its contraction), so buffer 1 may hold the scratch (extra_off) when smem is tight.
                float *wpart = reinterpret_cast<float *>(smem + a.extra_off);     // [NWARPS][TB][GNV]
                float *cpart = wpart + NWARPS * TB * GNV;                         // [MAX_CLUSTER][TB][GNV]
                float *fin = cpart + MAX_CLUSTER * TB * GNV;                      // [TB][GNV]
                float *Rp = fin + TB * GNV;                                       // [TB][RPK]
                float *gam = Rp + TB * RPK;                                       // [TB][GM + 1]
                unsigned *ctl = reinterpret_cast<unsigned *>(gam + TB * (GM + 1)); // [0] finished solves, [1] frozen in this cycle, [2] solves continued by the damped iteration
                float *fin2 = reinterpret_cast<float *>(ctl + 4);                 // [TB][GNV], second reduction of a step
                const bool warp_writer = (lane / KL == 0) && (kl % Own::SPLIT == 0);
                const bool cta_writer = warp == 0 && warp_writer;
                float4 *bas = a.basis + (size_t)(blockIdx.x / csize) * GM * TO4 * csize * nthreads + rank * nthreads + tid;
                const int bstride = csize * nthreads;
                float *wp_mine = wpart + (warp * TB + my_stim) * GNV;
                const float *fin_mine = fin + my_stim * GNV;

                // sum of the nv values per stimulus every warp has left in wpart -> fin (identical in every CTA)
                auto cluster_sum = [&](int nv, float *dst) {
                    __syncthreads();
                    const int b = tid / nv, i = tid - b * nv;
                    if (tid < TB * nv) {
                        float sacc = 0.f;
#pragma unroll
                        for (int w = 0; w < NWARPS; ++w) sacc += wpart[(w * TB + b) * GNV + i];
                        const unsigned dst = smem_u32(cpart + (rank * TB + b) * GNV + i);
                        for (int p = 0; p < csize; ++p) st_cluster_f32(map_to_rank(dst, p), sacc);
                    }
                    cluster.sync();
                    if (tid < TB * nv) {
                        float sacc = 0.f;
                        for (int p = 0; p < csize; ++p) sacc += cpart[(p * TB + b) * GNV + i];
                        dst[b * GNV + i] = sacc;
                    }
                    __syncthreads();
                };
                auto publish = [&](const float (&v)[TO]) {          // panel <- Phi v
#pragma unroll
                    for (int u = 0; u < TO; ++u)
                        if (valid[u]) {
                            const float af = phi[u] * v[u];
#pragma unroll
                            for (int p = 0; p < MAX_CLUSTER; ++p)
                                if (p < csize) st_cluster_f32(xpeer[p] + xoff[u], af);
                        }
                };
                auto store_vec = [&](int i, const float (&v)[TO]) {
#pragma unroll
                    for (int q = 0; q < TO4; ++q) {
                        float4 t4;
                        t4.x = v[4 * q];
                        t4.y = 4 * q + 1 < TO ? v[4 * q + 1 < TO ? 4 * q + 1 : 0] : 0.f;
                        t4.z = 4 * q + 2 < TO ? v[4 * q + 2 < TO ? 4 * q + 2 : 0] : 0.f;
                        t4.w = 4 * q + 3 < TO ? v[4 * q + 3 < TO ? 4 * q + 3 : 0] : 0.f;
                        __stcg(bas + (size_t)(i * TO4 + q) * bstride, t4);
                    }
                };
                auto load_vec = [&](int i, float (&v)[TO]) {
#pragma unroll
                    for (int q = 0; q < TO4; ++q) {
                        const float4 t4 = __ldcg(bas + (size_t)(i * TO4 + q) * bstride);
                        v[4 * q] = t4.x;
                        if (4 * q + 1 < TO) v[4 * q + 1 < TO ? 4 * q + 1 : 0] = t4.y;
                        if (4 * q + 2 < TO) v[4 * q + 2 < TO ? 4 * q + 2 : 0] = t4.z;
                        if (4 * q + 3 < TO) v[4 * q + 3 < TO ? 4 * q + 3 : 0] = t4.w;
                    }
                };

                // ---- ||g||^2 per stimulus: tolerance, and solves with nothing to do ----
                {
                    float p = 0.f;
#pragma unroll
                    for (int u = 0; u < TO; ++u) p = fmaf(g_own[u], g_own[u], p);
                    p = stim_lane_sum<KL>(p);
                    if (warp_writer) wp_mine[0] = p;
                }
                cluster_sum(1, fin);
                // A solve whose dL/dr vanishes (e.g. a rejected network masked out by the caller) has mu = 0: it is
                // finished before the first sweep and contributes nothing, whatever (possibly non-finite) state R holds.
                for (int b = 0; b < TB; ++b)
                    if (!(b < nact && fin[b * GNV] > 0.f)) dead |= 1u << b;
                if (tid == 0) { ctl[0] = dead; ctl[2] = 0u; }
                float beta = sqrtf(fmaxf(fin_mine[0], 0.f));
                const float tolabs = (float)a.rtol * beta;
                bool sdone = (dead >> my_stim) & 1u;
                float rres[TO];
#pragma unroll
                for (int u = 0; u < TO; ++u) { mu[u] = 0.0; rres[u] = g_own[u]; }
                if (sdone) {
                    my_status = 0;
#pragma unroll
                    for (int u = 0; u < TO; ++u) { phi[u] = 0.f; g_own[u] = 0.f; rres[u] = 0.f; }
                }
                int sweeps = 0, stagnant = 0;
                bool damped = false;                        // GMRES(16) stalled on this system: Richardson steps instead
                float best = 0.f;                           // smallest residual of the damped phase
                __syncthreads();                            // ctl[0]; fin is rewritten below
                while (ctl[0] != 0xffu) {
                    // ---------- one GMRES cycle from the true residual rres, |rres| = beta ----------
                    float cs[GM], sn[GM], vcur[TO];
                    float gcur = beta;
                    int kd = 0;
                    bool frozen = sdone || damped;
                    if (tid == 0) ctl[1] = ctl[0] | ctl[2];
                    {
                        const float inv = frozen ? 0.f : 1.f / beta;
#pragma unroll
                        for (int u = 0; u < TO; ++u) vcur[u] = rres[u] * inv;
                        store_vec(0, vcur);
                        publish(vcur);
                    }
#pragma unroll
                    for (int i = 0; i < GM; ++i) { cs[i] = 1.f; sn[i] = 0.f; }
                    // The panel holds Phi u_j with u_j = v_j / scale: the new direction is published BEFORE its norm is
                    // known (the publish shares the cluster barrier of the norm's reduction) and the contraction's
                    // result is scaled afterwards -- two cluster barriers per Arnoldi step instead of three.
                    float scale = 1.f;
                    cluster.sync();
                    for (int j = 0; j < GM && ctl[1] != 0xffu; ++j) {
                        float wv[TO];
                        {
                            float acc[TI][TB], y[TO];
                            contract_panel<TI, KL>(acc, Wsm, X4, P, kpad, 0, wrow, kl);
                            reduce_scatter<TI, KL>(acc, y, kl);
#pragma unroll
                            for (int u = 0; u < TO; ++u) wv[u] = valid[u] ? fmaf(-scale, y[u], vcur[u]) : 0.f;   // w = A v_j
                        }
                        ++sweeps;
                        // classical Gram-Schmidt: h_i = <w, v_i> (i <= j), one cluster reduction ...
                        // (the basis vectors v_0..v_j come from the L2 scratch: all loads of a step are issued together and
                        // kept in registers for the orthogonalisation below -- one L2 latency per step instead of 2 (j + 1))
                        constexpr bool CACHE = TO <= 4;
                        float vb[CACHE ? GM : 1][TO];
                        if constexpr (CACHE) {
#pragma unroll
                            for (int i = 0; i < GM; ++i)
                                if (i <= j) load_vec(i, vb[i]);
#pragma unroll
                            for (int i = 0; i < GM; ++i)
                                if (i <= j) {
                                    float p = 0.f;
#pragma unroll
                                    for (int u = 0; u < TO; ++u) p = fmaf(wv[u], vb[i][u], p);
                                    p = stim_lane_sum<KL>(p);
                                    if (warp_writer) wp_mine[i] = p;
                                }
                        } else {
#pragma unroll 4
                            for (int i = 0; i <= j; ++i) {
                                float vi[TO], p = 0.f;
                                load_vec(i, vi);
#pragma unroll
                                for (int u = 0; u < TO; ++u) p = fmaf(wv[u], vi[u], p);
                                p = stim_lane_sum<KL>(p);
                                if (warp_writer) wp_mine[i] = p;
                            }
                        }
                        cluster_sum(j + 1, fin);
                        // ... w' = w - sum h_i v_i, and a second (one value) reduction for |w'|^2.  Taking the norm from
                        // |w|^2 - sum h_i^2 instead would save this barrier, but it is unstable here: the systems are
                        // I - (small), so h_jj ~ 1 >> |w'|, a 0.1 % error in the norm of v_j becomes a 50 % error in
                        // |w'|^2 one step later and the basis blows up within a cycle (measured: estimated residuals off
                        // by 1e30, cycles that increase the true residual; with the true norm the ratio stays below 7
                        // and the solves need ~12 % fewer contractions).
                        float vnext[TO], hsq = 0.f;
#pragma unroll
                        for (int u = 0; u < TO; ++u) vnext[u] = wv[u];
                        if (!frozen) {
                            if constexpr (CACHE) {
#pragma unroll
                                for (int i = 0; i < GM; ++i)
                                    if (i <= j) {
                                        const float h = fin_mine[i];
                                        hsq = fmaf(h, h, hsq);
#pragma unroll
                                        for (int u = 0; u < TO; ++u) vnext[u] = fmaf(-h, vb[i][u], vnext[u]);
                                    }
                            } else {
#pragma unroll 4
                                for (int i = 0; i <= j; ++i) {
                                    float vi[TO];
                                    load_vec(i, vi);
                                    const float h = fin_mine[i];
                                    hsq = fmaf(h, h, hsq);
#pragma unroll
                                    for (int u = 0; u < TO; ++u) vnext[u] = fmaf(-h, vi[u], vnext[u]);
                                }
                            }
                        }
                        {
                            float p = 0.f;
#pragma unroll
                            for (int u = 0; u < TO; ++u) p = fmaf(vnext[u], vnext[u], p);
                            p = stim_lane_sum<KL>(p);
                            if (warp_writer) wp_mine[0] = p;
                        }
                        publish(vnext);                 // unnormalised; every CTA is past its contraction (barrier above)
                        cluster_sum(1, fin2);
                        if (!frozen) {
                            // new column of the Hessenberg matrix, rotated into R (every thread of the stimulus
                            // computes the same numbers; one of them records R and the rotated right-hand side)
                            float col[GM + 1];
#pragma unroll
                            for (int i = 0; i < GM; ++i) col[i] = i <= j ? fin_mine[i] : 0.f;
                            const float hn2 = fin2[my_stim * GNV];                            // |w - sum h_i v_i|^2
                            const bool brk = !(hn2 > 1e-9f * (hsq + hn2));    // Krylov space exhausted (or NaN)
                            const float hn = sqrtf(fmaxf(hn2, 0.f));
                            float gnext = 0.f;
#pragma unroll
                            for (int i = 0; i < GM; ++i) {
                                if (i < j) {
                                    const float t0 = col[i], t1 = col[i + 1];
                                    col[i] = fmaf(cs[i], t0, sn[i] * t1);
                                    col[i + 1] = fmaf(cs[i], t1, -sn[i] * t0);
                                } else if (i == j) {
                                    const float d = sqrtf(fmaf(col[i], col[i], hn * hn));
                                    const float c = d > 0.f ? col[i] / d : 1.f, sg = d > 0.f ? hn / d : 0.f;
                                    cs[i] = c; sn[i] = sg;
                                    col[i] = d;
                                    gnext = -sg * gcur;
                                    gcur = c * gcur;
                                }
                            }
                            if (cta_writer) {
                                float *Rc = Rp + my_stim * RPK + j * (j + 1) / 2;
#pragma unroll
                                for (int i = 0; i < GM; ++i)
                                    if (i <= j) Rc[i] = col[i];
                                gam[my_stim * (GM + 1) + j] = gcur;
                            }
                            gcur = gnext;
                            kd = j + 1;
                            if (brk || fabsf(gnext) <= 0.7f * tolabs || j == GM - 1) {
                                frozen = true;
                                scale = 0.f;
                                if (cta_writer) atomicOr(&ctl[1], 1u << my_stim);
#pragma unroll
                                for (int u = 0; u < TO; ++u) vnext[u] = 0.f;
                            } else {
                                scale = 1.f / hn;
#pragma unroll
                                for (int u = 0; u < TO; ++u) vnext[u] *= scale;
                            }
                        } else {
                            scale = 0.f;
#pragma unroll
                            for (int u = 0; u < TO; ++u) vnext[u] = 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < TO; ++u) vcur[u] = vnext[u];
                        if (j + 1 < GM) store_vec(j + 1, vcur);
                        __syncthreads();                // ctl[1], R and the rotated right-hand side of this step
                        if (ctl[1] == 0xffu) break;
                    }
                    // ---------- mu += V y,  R y = gamma ----------
                    if (!sdone && damped) {
                        // one step of mu <- mu + eps (g - mu + W^T Phi mu): contracts wherever the forward Euler scheme does
#pragma unroll
                        for (int u = 0; u < TO; ++u) mu[u] += eps_own[u] * (double)rres[u];
                    } else if (!sdone && kd > 0) {
                        float yv[GM];
                        const float *Rs = Rp + my_stim * RPK, *gs = gam + my_stim * (GM + 1);
#pragma unroll
                        for (int i = 0; i < GM; ++i) yv[i] = i < kd ? gs[i] : 0.f;
#pragma unroll
                        for (int i = GM - 1; i >= 0; --i)
                            if (i < kd) {
                                yv[i] = yv[i] / Rs[i * (i + 1) / 2 + i];
#pragma unroll
                                for (int k2 = 0; k2 < i; ++k2) yv[k2] = fmaf(-Rs[i * (i + 1) / 2 + k2], yv[i], yv[k2]);
                            }
#pragma unroll
                        for (int i = 0; i < GM; ++i)
                            if (i < kd) {
                                float vi[TO];
                                load_vec(i, vi);
#pragma unroll
                                for (int u = 0; u < TO; ++u) mu[u] += (double)yv[i] * (double)vi[u];
                            }
                    }
                    // ---------- true residual g - mu + W^T Phi mu ----------
                    {
                        float pm[TO];
#pragma unroll
                        for (int u = 0; u < TO; ++u) pm[u] = (float)mu[u];
                        publish(pm);
                    }
                    cluster.sync();
                    {
                        float acc[TI][TB], y[TO];
                        contract_panel<TI, KL>(acc, Wsm, X4, P, kpad, 0, wrow, kl);
                        reduce_scatter<TI, KL>(acc, y, kl);
                        float p = 0.f;
#pragma unroll
                        for (int u = 0; u < TO; ++u) {
                            rres[u] = (valid[u] && !sdone) ? (float)((double)g_own[u] - mu[u] + (double)y[u]) : 0.f;
                            p = fmaf(rres[u], rres[u], p);
                        }
                        p = stim_lane_sum<KL>(p);
                        if (warp_writer) wp_mine[0] = p;
                    }
                    ++sweeps;
                    cluster_sum(1, fin);
                    if (!sdone) {
                        const float bnew = sqrtf(fmaxf(fin_mine[0], 0.f));
                        my_iters = sweeps;
                        if (bnew <= tolabs) { sdone = true; my_status = 0; }
                        else if (!(bnew == bnew) || sweeps >= a.max_iter) sdone = true;
                        else if (!damped) {
                            stagnant = bnew < 0.9f * beta ? 0 : stagnant + 1;
                            if (a.test_stall && bnew > 16.f * tolabs) stagnant = 2;
                            if (stagnant >= 2) {
                                // Two cycles without progress.  Within 16 x the tolerance this is the floor of the FP32
                                // residual (~1e-7 |W^T Phi mu|): accepted.  Otherwise restarted GMRES has stalled on this
                                // system and the solve continues with the damped iteration, one contraction per step.
                                if (bnew <= 16.f * tolabs) { sdone = true; my_status = 0; }
                                else {
                                    damped = true; stagnant = 0; best = bnew;
                                    if (cta_writer) atomicOr(&ctl[2], 1u << my_stim);
                                }
                            }
                        } else {
                            // the residual norm of the damped iteration need not fall monotonically (W^T Phi is far
                            // from normal): give up only after 64 steps without a new best
                            if (bnew < 0.99f * best) { best = bnew; stagnant = 0; }
                            else if (++stagnant >= 64) { sdone = true; if (bnew <= 16.f * tolabs) my_status = 0; }
                        }
                        beta = bnew;
                        if (sdone && cta_writer) atomicOr(&ctl[0], 1u << my_stim);
                    }
                    __syncthreads();
                }
                // The panel holds Phi mu of every stimulus, published for the last true residual -- unless no cycle ran
                // at all; stimuli with nothing to do must not leave their (possibly non-finite) state r behind.
                if (dead == 0xffu) {                           // uniform over the cluster
                    float pm[TO];
#pragma unroll
                    for (int u = 0; u < TO; ++u) pm[u] = 0.f;
                    publish(pm);
                    cluster.sync();
                }
            } else {
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
                my_iters = a.max_iter;
                // A solve whose dL/dr vanishes (e.g. a rejected network masked out by the caller) has mu = 0: it is
                // finished before the first sweep and contributes nothing, whatever (possibly non-finite) state R holds.
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
                buf = 1;
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
#pragma unroll 8
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
struct IftVariant { IftKernel gmres, damped; int threads; int rows; int kl; int to4; };
// rows covered = (32/KL) * NWARPS * TI; kpad must be a multiple of 4*KL (same table as the forward fallback kernel)
static const IftVariant kIftVariants[] = {
    {ssn_ift_cluster_kernel<4, 16, 8, true>, ssn_ift_cluster_kernel<4, 16, 8, false>, 256, 64, 16, 1},
    {ssn_ift_cluster_kernel<7, 16, 8, true>, ssn_ift_cluster_kernel<7, 16, 8, false>, 256, 112, 16, 1},
    {ssn_ift_cluster_kernel<7, 8, 8, true>, ssn_ift_cluster_kernel<7, 8, 8, false>, 256, 224, 8, 2},
};

// Smallest cluster whose CTAs hold their slice of W^T plus the solver's scratch.  The GMRES scratch goes behind the
// regular layout when there is room, otherwise into the second panel buffer (which that solver does not use).
static bool choose_ift_shape(int n_sites, int smem_limit, bool gmres, ClusterShape *out, int *variant, int *extra_off,
                             int *smem_total) {
    const int dim = 2 * n_sites;
    ClusterShape s;
    s.dim = dim;
    for (int c = 1; c <= MAX_CLUSTER; c *= 2) {
        s.csize = c;
        s.rpc = (dim + c - 1) / c;
        if (s.rpc * (c - 1) >= dim) continue;               // a CTA would own no rows
        for (int v = 0; v < 3; ++v) {
            if (kIftVariants[v].rows < s.rpc) continue;
            const int q = 4 * kIftVariants[v].kl;
            s.kpad = ((dim + q - 1) / q) * q;
            const SmemLayout L = smem_layout(s, n_sites);
            const int buf_bytes = 2 * panel_P(s.kpad) * 16;
            int total = L.total, off = 0;
            if (gmres) {
                if (L.total + IFT_EXTRA_SMEM <= smem_limit) { off = L.total; total = L.total + IFT_EXTRA_SMEM; }
                else if (buf_bytes >= IFT_EXTRA_SMEM) off = L.x_off + buf_bytes;
                else total = smem_limit + 1;
            }
            if (total <= smem_limit) { *out = s; *variant = v; *extra_off = off; *smem_total = total; return true; }
            break;
        }
    }
    return false;
}

static bool ift_use_damped() {
    const char *e = getenv("SSN_IFT");
    return e && !strcmp(e, "damped");
}

int launch_ift_gradient(const ssn_solver &sv, int nz, int nb, int n_sites, const float *z, const ssn_jds &jds,
                        const float *ext, int ext_per_network, const float *R, const float *g, double rtol,
                        double *grad, float *mu, int *status, int *iters, float *grad_ext, int *counter,
                        cudaStream_t stream) {
    SSN_CUDA(cudaMemsetAsync(grad, 0, 12 * sizeof(double), stream));
    if (nz <= 0 || nb <= 0) return 0;
    int dev = 0, limit = 0, variant = 0, smem = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const bool gmres = !ift_use_damped();
    IftArgs a = {};
    if (!choose_ift_shape(n_sites, limit - 256, gmres, &a.shape, &variant, &a.extra_off, &smem)) {
        set_error("ift kernel: 2N=%d does not fit a cluster of %d CTAs", 2 * n_sites, MAX_CLUSTER);
        return -1;
    }
    const IftVariant var = kIftVariants[variant];
    const IftKernel fn = gmres ? var.gmres : var.damped;
    SSN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    a.nz = nz; a.nb = nb; a.n_sites = n_sites;
    a.z = z; a.wc = make_weight_const(jds, n_sites);
    a.ext = ext; a.ext_stride_z = ext_per_network ? (long long)nb * 2 * n_sites : 0;
    a.R = R; a.g = g; a.mu = mu; a.grad_ext = grad_ext; a.status = status; a.iters = iters; a.grad = grad; a.work_counter = counter;
    a.io = make_io_const<float>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.eps_E = sv.dt / sv.tau_E; a.eps_I = sv.dt / sv.tau_I;
    a.rtol = rtol > 0 ? rtol : 1e-6;
    a.max_iter = sv.max_iter;
    {
        const char *e = getenv("SSN_IFT");
        a.test_stall = e && !strcmp(e, "stall");
    }

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
    SSN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg));
    if (max_clusters < 1) { set_error("ift kernel: no resident cluster"); return -1; }
    const int clusters = std::min(max_clusters, nz);
    cfg.gridDim = dim3(clusters * a.shape.csize, 1, 1);
    SSN_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    if (gmres) {
        // Krylov basis: stream-ordered scratch (one launch's worth: 10 MB at 2N = 402, it stays in L2), so that
        // concurrent launches on other streams never share it
        const size_t bytes = (size_t)clusters * GM * var.to4 * a.shape.csize * var.threads * sizeof(float4);
        static std::atomic<unsigned long long> pool_ready{0};                 // bit per device: keep freed blocks cached
        if (!(pool_ready.load() >> (dev & 63) & 1ull)) {
            cudaMemPool_t pool;
            unsigned long long keep = ~0ull;
            SSN_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            SSN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            pool_ready.fetch_or(1ull << (dev & 63));
        }
        SSN_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&a.basis), bytes, stream));
    }
    cudaError_t launched;
    {
        KernelTimer kt("ssn_ift_cluster_kernel", stream);
        launched = cudaLaunchKernelEx(&cfg, fn, a);
    }
    if (gmres) SSN_CUDA(cudaFreeAsync(a.basis, stream));
    SSN_CUDA(launched);
    count_launch();
    return 0;
}

}  // namespace ssn
