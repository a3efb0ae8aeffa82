// K3 / K4: fixed-length unrolled Euler dynamics of the SSN and BPTT through it.
//
// Replaces the Theano graph of tc_gan/networks/ssn.py: EulerSSNCore.get_output_for
// (:555-576) unrolled by Lasagne's CustomRecurrentLayer (:354-385), the outputs of
// EulerSSNModel (:619-633: time_avg, dynamics_penalty, rate_penalty) and Theano's
// autodiff through the scan.
//
//   forward   r_0 = 0,  r_{k+1} = (1-eps) r_k + eps f(W r_k + I),  k = 0..seqlen-1
//             traj[t] = r_{t+1};  gain[t] = eps f'(W r_t + I)
//   backward  lambda_k = d_k + (1-eps) lambda_{k+1} + W^T q_k,  q_k = gain[k] * lambda_{k+1}
//             dL/dW = sum_k q_k r_k^T   (a [2N x K] x [K x 2N] contraction, K = seqlen * nb)
//             dL/dtheta = <dL/dW, dW/dtheta>  fused into the contraction's epilogue
//
// Forward and the adjoint recursion run on the cluster-resident machinery of K1/K2
// (W, then W^T, in distributed shared memory; one DSMEM exchange per time step).
//
// Exchange: no cluster barrier inside the time loop.  A thread publishes its outputs of step t with st.async into
// panel buffer (t+1)&1 of every CTA of the cluster (its own included); the stores complete bytes on that buffer's
// mbarrier in the destination CTA, and a CTA starts step t+1 when 2N x 8 x 4 bytes have landed.  Double buffering
// is enough: a peer can publish step t+1 only after its step-t+1 contraction, which needs every output of step t
// of this CTA, i.e. all of this CTA's warps are past their reads of the buffer being overwritten.
#include <cstdlib>
#include <cstring>
#include "ssn_ws_common.cuh"
#include "ssn_launch.h"

namespace ssn {

bool choose_cluster_shape(int n_sites, ClusterShape *out, int smem_limit, int *variant);

struct EulerArgs {
    int nz, nb, n_sites;
    ClusterShape shape;
    const float *z;
    WeightConst wc;
    const float *ext;
    long long ext_stride_z;
    int seqlen, skip;
    int pitch;                          // floats per (time step, stimulus) row of traj / gain / adj: 2N rounded up to 4
    float threshold;
    int *work_counter;
    IoConst<float> io;
    double eps_E, eps_I;
    // forward
    float *time_avg;
    double *penalties;
    float *traj, *gain;
    // backward
    const float *g_avg;
    double w_dyn, w_rate;
    const float *w_dev;                // optional device multipliers of (w_dyn, w_rate)
    const float *traj_in, *gain_in;
    float *adj, *grad_ext;
};

template <int TI, int KL, int NWARPS, bool BACKWARD>
__global__ void __launch_bounds__(NWARPS * 32, 1) ssn_euler_cluster_kernel(const EulerArgs a) {
    using Own = Owner<TI, KL>;
    constexpr int TO = Own::TO;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[2];
    __shared__ __align__(8) unsigned long long bars[2];      // "panel buffer b is complete", one phase per use
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int csize = a.shape.csize;
    const int dim = a.shape.dim, kpad = a.shape.kpad, rpc = a.shape.rpc, N = a.n_sites;
    const int P = panel_P(kpad);
    const SmemLayout L = smem_layout(a.shape, a.n_sites);
    float *Wsm = reinterpret_cast<float *>(smem + L.w_off);
    float *Xf = reinterpret_cast<float *>(smem + L.x_off);
    const float4 *X4 = reinterpret_cast<const float4 *>(smem + L.x_off);
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

    unsigned xpeer[MAX_CLUSTER], bpeer[MAX_CLUSTER];
#pragma unroll
    for (int p = 0; p < MAX_CLUSTER; ++p) {
        xpeer[p] = map_to_rank(smem_u32(Xf), p < csize ? p : 0);
        bpeer[p] = map_to_rank(smem_u32(&bars[0]), p < csize ? p : 0);
    }

    build_profile_table(a.wc, N, gtab, tid, nthreads);
    for (int i = tid; i < 2 * 2 * P * 4; i += nthreads) Xf[i] = 0.f;
    if (tid < 2) red[tid] = 0.0;
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_chunks = (a.nb + TB - 1) / TB;
    const unsigned buf_bytes = 2u * (unsigned)P * 16u;
    const unsigned step_bytes = (unsigned)dim * TB * 4u;       // every (row, stimulus slot) of the panel, once per step
    const int seqlen = a.seqlen, skip = a.skip;
    const int T = seqlen - skip;
    double pen_dyn = 0.0, pen_rate = 0.0;
    unsigned phase = 0u;                                        // bit b: parity of the next completion of bars[b]
    // publish one value per owned row into buffer nb of every CTA ...
    auto publish = [&](const float (&val)[TO], const unsigned (&xo)[TO], int nb) {
        const unsigned boff = nb ? buf_bytes : 0u;
#pragma unroll
        for (int u = 0; u < TO; ++u)
            if (valid[u]) {
#pragma unroll
                for (int p = 0; p < MAX_CLUSTER; ++p)
                    if (p < csize) st_async_u32(xpeer[p] + xo[u] + boff, __float_as_uint(val[u]), bpeer[p] + 8u * nb);
            }
        if (tid == 0) mbar_arrive_expect_tx(smem_u32(&bars[nb]), step_bytes);
    };
    // ... and wait for the whole panel of that buffer (after whatever work does not feed the peers)
    auto await = [&](int nb) {
        mbar_wait(smem_u32(&bars[nb]), (phase >> nb) & 1u);
        phase ^= 1u << nb;
    };
    for (;;) {
        if (rank == 0 && tid == 0) {
            const int n = atomicAdd(a.work_counter, 1);
            for (int p = 0; p < csize; ++p) st_cluster_u32(map_to_rank(smem_u32(&misc->next_net), p), (unsigned)n);
        }
        cluster.sync();
        const int net = misc->next_net;
        if (net >= a.nz) break;
        const float *z_net = a.z + (size_t)net * dim * dim;

        load_matrix_slice<BACKWARD>(Wsm, z_net, SSN_W_FROM_Z, a.wc, gtab, N, dim, kpad, row_base, rows_here,
                                    tid, nthreads);

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int b0 = chunk * TB;
            const int nact = min(TB, a.nb - b0);
            const bool active = my_stim < nact;
            const int stim_g = b0 + my_stim;

            unsigned xoff[TO];
            double eps_own[TO];
            size_t goff[TO];                           // offset of (stim, row) inside one [nb][2N] slice
            size_t toff[TO];                           // ... inside one [nb][pitch] time slice of traj / gain / adj
#pragma unroll
            for (int u = 0; u < TO; ++u) {
                xoff[u] = 0u; eps_own[u] = 0.0; goff[u] = 0; toff[u] = 0;
                if (valid[u]) {
                    const int gr = row_base + own0 + u;
                    xoff[u] = 4u * (unsigned)panel_index(P, 0, gr, my_stim);
                    eps_own[u] = gr < N ? a.eps_E : a.eps_I;
                    goff[u] = (size_t)stim_g * dim + gr;
                    toff[u] = (size_t)stim_g * a.pitch + gr;
#pragma unroll
                    for (int p = 0; p < MAX_CLUSTER; ++p)
                        if (p < csize) st_cluster_f32(xpeer[p] + xoff[u], 0.f);
                }
            }
            const size_t slice = (size_t)a.nb * dim;                    // one [nb][2N] array of one network
            const size_t tslice = (size_t)a.nb * a.pitch;               // one time step of one network in traj / gain / adj
            const size_t net_base = (size_t)net * seqlen * tslice;

            if (!BACKWARD) {
                float ext_own[TO];
                double avg[TO];
#pragma unroll
                for (int u = 0; u < TO; ++u) {
                    avg[u] = 0.0;
                    ext_own[u] = (valid[u] && active)
                        ? __ldg(a.ext + (size_t)net * a.ext_stride_z + goff[u]) : 0.f;
                }
                cluster.sync();
                int buf = 0;
                float eps_f[TO], rstate[TO];
#pragma unroll
                for (int u = 0; u < TO; ++u) { eps_f[u] = (float)eps_own[u]; rstate[u] = 0.f; }
                const bool pure_power = a.io.io_type == SSN_IO_POWER;
                for (int t = 0; t < seqlen; ++t) {
                    float acc[TI][TB], v[TO], rpub[TO];
                    contract_panel<TI, KL>(acc, Wsm, X4, P, kpad, buf, wrow, kl);
                    reduce_scatter<TI, KL>(acc, v, kl);
                    const int nbuf = buf ^ 1;
                    const bool keep = t >= skip;
                    float r_prev[TO], gv[TO];
#pragma unroll
                    for (int u = 0; u < TO; ++u) {
                        const float vt = v[u] + ext_own[u];
                        // power-law branch without a branch (the common case); the saturating branch of asym_tanh /
                        // asym_linear only for the lanes above v0
                        float fv;
                        io_power_fast(a.io, fmaxf(vt, 1e-30f), fv, gv[u]);
                        fv = vt > 0.f ? fv : 0.f;
                        gv[u] = vt > 0.f ? gv[u] : 0.f;
                        if (!pure_power && vt > a.io.v0) { fv = io_eval<float>(a.io, vt); gv[u] = io_gain<float>(a.io, vt); }
                        r_prev[u] = rstate[u];
                        rstate[u] = fmaf(eps_f[u], fv - rstate[u], rstate[u]);      // float32 state, as the reference (floatX)
                        rpub[u] = rstate[u];
                    }
                    publish(rpub, xoff, nbuf);          // the peers wait for this: outputs to HBM and the penalties come after
#pragma unroll
                    for (int u = 0; u < TO; ++u)
                        if (valid[u] && active) {
                            if (keep) {
                                const double r_new = (double)rpub[u];
                                avg[u] += r_new;
                                pen_rate += fmax(r_new - (double)a.threshold, 0.0);
                                if (t > skip) pen_dyn += (r_new - (double)r_prev[u]) * (r_new - (double)r_prev[u]);
                            }
                            const size_t o = net_base + (size_t)t * tslice + toff[u];
                            if (a.traj) a.traj[o] = rpub[u];
                            if (a.gain) a.gain[o] = eps_f[u] * gv[u];
                        }
                    await(nbuf);
                    buf = nbuf;
                }
#pragma unroll
                for (int u = 0; u < TO; ++u)
                    if (valid[u] && active) a.time_avg[(size_t)net * slice + goff[u]] = (float)(avg[u] / T);
            } else {
                // ---- adjoint recursion, k = seqlen .. 1 (array index tp = k - 1) ----
                const double w_dyn = a.w_dev ? a.w_dyn * (double)__ldg(a.w_dev) : a.w_dyn;
                const double w_rate = a.w_dev ? a.w_rate * (double)__ldg(a.w_dev + 1) : a.w_rate;
                const float w_rate_f = (float)w_rate, w_dyn2_f = (float)(2.0 * w_dyn);
                float gavg[TO], r_cur[TO], r_next[TO], gext[TO], lstate[TO], omeps_f[TO];
#pragma unroll
                for (int u = 0; u < TO; ++u) {
                    gavg[u] = 0.f; r_cur[u] = 0.f; r_next[u] = 0.f; gext[u] = 0.f; lstate[u] = 0.f;
                    omeps_f[u] = (float)(1.0 - eps_own[u]);
                    if (valid[u] && active) {
                        gavg[u] = __ldg(a.g_avg + (size_t)net * slice + goff[u]) / (float)T;
                        r_cur[u] = __ldg(a.traj_in + net_base + (size_t)(seqlen - 1) * tslice + toff[u]);
                    }
                }
                cluster.sync();
                int buf = 0;
                for (int tp = seqlen - 1; tp >= 0; --tp) {
                    const int k = tp + 1;
                    // operands of this step from HBM, requested before the contraction that hides their latency
                    float gain_k[TO], r_prevs[TO];
#pragma unroll
                    for (int u = 0; u < TO; ++u) {
                        const bool ld = valid[u] && active;
                        gain_k[u] = ld ? __ldg(a.gain_in + net_base + (size_t)tp * tslice + toff[u]) : 0.f;
                        r_prevs[u] = (ld && tp > 0) ? __ldg(a.traj_in + net_base + (size_t)(tp - 1) * tslice + toff[u]) : 0.f;
                    }
                    float y[TO], qpub[TO];
#pragma unroll
                    for (int u = 0; u < TO; ++u) y[u] = 0.f;
                    if (k < seqlen) {                   // q_k was published at the end of the previous step
                        float acc[TI][TB];
                        contract_panel<TI, KL>(acc, Wsm, X4, P, kpad, buf, wrow, kl);
                        reduce_scatter<TI, KL>(acc, y, kl);
                    }
                    const int nbuf = buf ^ 1;
#pragma unroll
                    for (int u = 0; u < TO; ++u) {
                        qpub[u] = 0.f;
                        if (valid[u]) {
                            const float r_prev = r_prevs[u];
                            float lam = 0.f;
                            if (active) {
                                float d = 0.f;
                                if (k >= skip + 1) {
                                    d = gavg[u];
                                    if (r_cur[u] > a.threshold) d += w_rate_f;
                                    // d/dr_k of sum (r_{t+1} - r_t)^2: the two differences first, then one scaling
                                    float dd = 0.f;
                                    if (k >= skip + 2) dd = r_cur[u] - r_prev;
                                    if (k <= seqlen - 1) dd -= r_next[u] - r_cur[u];
                                    d = fmaf(w_dyn2_f, dd, d);
                                }
                                lam = fmaf(omeps_f[u], lstate[u], d) + y[u];
                            }
                            lstate[u] = lam;                          // lambda_k (float32, like the reference's floatX graph)
                            // q_{k-1} = gain[k-1] * lambda_k, paired with r_{k-1} = traj[tp-1]
                            float q = 0.f;
                            if (active) {
                                q = gain_k[u] * lam;
                                gext[u] += q;                             // dL/d ext = sum_k q_k, q_0 included
                                if (tp == 0) q = 0.f;                     // q_0 pairs with r_0 = 0: nothing to publish
                            }
                            r_next[u] = r_cur[u];
                            r_cur[u] = r_prev;
                            qpub[u] = q;
                        }
                    }
                    publish(qpub, xoff, nbuf);          // the peers wait for this: the stores to HBM come after
#pragma unroll
                    for (int u = 0; u < TO; ++u)
                        if (valid[u] && active) {
                            if (tp > 0) a.adj[net_base + (size_t)(tp - 1) * tslice + toff[u]] = qpub[u];
                            if (tp == seqlen - 1) a.adj[net_base + (size_t)tp * tslice + toff[u]] = 0.f;   // q_seqlen = 0
                        }
                    await(nbuf);
                    buf = nbuf;
                }
                if (a.grad_ext) {
#pragma unroll
                    for (int u = 0; u < TO; ++u)
                        if (valid[u] && active) a.grad_ext[(size_t)net * slice + goff[u]] = gext[u];
                }
            }
            cluster.sync();
        }
    }
    if (!BACKWARD) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pen_dyn += __shfl_xor_sync(0xffffffffu, pen_dyn, o);
            pen_rate += __shfl_xor_sync(0xffffffffu, pen_rate, o);
        }
        if (lane == 0) { atomicAdd(&red[0], pen_dyn); atomicAdd(&red[1], pen_rate); }
        __syncthreads();
        if (tid < 2) atomicAdd(a.penalties + tid, red[tid]);
    }
}

// ------------------------------------------------------------------------------------
// dL/dtheta = < sum_k q_k r_k^T , dW/dtheta >:  per network a [dim x K] x [K x dim]
// contraction (K = seqlen * nb) tiled 64 x 64 per CTA, FP32 FFMA, with the product
// against dW/dtheta (z re-read, Gaussian profile recomputed) fused into the epilogue.
// ------------------------------------------------------------------------------------
constexpr int GT = 128, GK = 8;          // 128 x 128 output tile per CTA, K chunks of 8, 8 x 8 outputs per thread

__global__ void __launch_bounds__(256) ssn_bptt_param_grad_kernel(int n_sites, int pitch, long long K, const float *adj,
                                                                  const float *traj, const float *z,
                                                                  WeightConst wc, double *grad) {
    // double-buffered K chunks: [2][GK][GT] for each operand (2 x 2 x 8 x 128 x 4 B = 16 KB)
    __shared__ __align__(16) float As[2][GK][GT], Bs[2][GK][GT];
    __shared__ double red[12];
    const int dim = 2 * n_sites;
    const int tiles = (dim + GT - 1) / GT;
    // effective tile edge: the smallest multiple of 8 that covers dim with `tiles` tiles (104 for 2N = 402,
    // so 93 % of the FMAs are useful instead of 62 % with 128)
    const int gt = (((dim + tiles - 1) / tiles + 7) / 8) * 8, half = gt / 2, n1 = gt / 8;
    const int net = blockIdx.x / (tiles * tiles);
    const int ti = (blockIdx.x / tiles) % tiles, tj = blockIdx.x % tiles;
    const int i0 = ti * gt, j0 = tj * gt;
    const float *A = adj + (size_t)net * K * pitch, *B = traj + (size_t)net * K * pitch;
    const int tid = threadIdx.x;
    const bool worker = tid < n1 * n1;                      // n1 x n1 threads, 8 x 8 outputs each
    const int tx = worker ? tid % n1 : 0, ty = worker ? tid / n1 : 0;   // rows ty*4 + {0..3, half..half+3}, cols likewise
    if (tid < 12) red[tid] = 0.0;
    float c[8][8];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) c[p][q] = 0.f;
    // loader: 256 threads x 4 elements = GK x GT per operand per chunk (scalar, rows of adj/traj are only 8-byte aligned)
    const int lk = tid / 32, lc = tid % 32;                 // k row lk (0..7), columns lc + 32 m
    const long long nchunks = (K + GK - 1) / GK;
    float ra[4], rb[4];
    auto fetch = [&](long long chunk) {
        const long long k = chunk * GK + lk;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int col = lc + 32 * m;
            ra[m] = (k < K && col < gt && i0 + col < dim) ? __ldg(A + k * pitch + i0 + col) : 0.f;
            rb[m] = (k < K && col < gt && j0 + col < dim) ? __ldg(B + k * pitch + j0 + col) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            As[buf][lk][lc + 32 * m] = ra[m];
            Bs[buf][lk][lc + 32 * m] = rb[m];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (long long ch = 0; ch < nchunks; ++ch) {
        const int buf = (int)(ch & 1);
        if (ch + 1 < nchunks) fetch(ch + 1);                 // global loads in flight during the FMAs
        if (worker)
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][half + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][half + tx * 4]);
            const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int q = 0; q < 8; ++q) c[p][q] = fmaf(aa[p], bb[q], c[p][q]);
        }
        if (ch + 1 < nchunks) stash(buf ^ 1);
        __syncthreads();
    }
    // fused epilogue: <dL/dW tile, dW/dtheta> with z re-read and the Gaussian profile recomputed
    const float *z_net = z + (size_t)net * dim * dim;
    float sJ[4] = {0.f, 0.f, 0.f, 0.f}, sD[4] = {0.f, 0.f, 0.f, 0.f}, sS[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + (p < 4 ? ty * 4 + p : half + ty * 4 + p - 4);
            const int j = j0 + (q < 4 ? tx * 4 + q : half + tx * 4 + q - 4);
            if (worker && i < dim && j < dim) {
                const int ah = i >= n_sites, bh = j >= n_sites, ab = ah * 2 + bh;
                const float zz = __ldg(z_net + (size_t)i * dim + j);
                const float x = (float)((i - ah * n_sites) - (j - bh * n_sites)) * wc.dx;
                const float gG = expf(-x * x * wc.inv2s2[ab]) * c[p][q];
                const float sgn = bh == 0 ? 1.f : -1.f;
                const float vJ = sgn * gG, vD = sgn * gG * zz;
                const float vS = gG * x * x * wc.invS3[ab] * fmaf(wc.sD[ab], zz, wc.sJ[ab]);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    sJ[w] += w == ab ? vJ : 0.f;
                    sD[w] += w == ab ? vD : 0.f;
                    sS[w] += w == ab ? vS : 0.f;
                }
            }
        }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sJ[w] += __shfl_xor_sync(0xffffffffu, sJ[w], o);
            sD[w] += __shfl_xor_sync(0xffffffffu, sD[w], o);
            sS[w] += __shfl_xor_sync(0xffffffffu, sS[w], o);
        }
        if ((tid & 31) == 0) {
            atomicAdd(&red[w], (double)sJ[w]);
            atomicAdd(&red[4 + w], (double)sD[w]);
            atomicAdd(&red[8 + w], (double)sS[w]);
        }
    }
    __syncthreads();
    if (tid < 12 && red[tid] != 0.0) atomicAdd(grad + tid, red[tid]);
}

// ------------------------------------------------------------------------------------

typedef void (*EulerKernel)(const EulerArgs);
struct EulerVariant { EulerKernel fwd, bwd; int threads; };
static const EulerVariant kEulerVariants[] = {
    {ssn_euler_cluster_kernel<4, 16, 8, false>, ssn_euler_cluster_kernel<4, 16, 8, true>, 256},
    {ssn_euler_cluster_kernel<7, 16, 8, false>, ssn_euler_cluster_kernel<7, 16, 8, true>, 256},
    {ssn_euler_cluster_kernel<7, 8, 8, false>, ssn_euler_cluster_kernel<7, 8, 8, true>, 256},
};

static int launch_euler(bool backward, EulerArgs &a, int n_sites, int nz, int *counter, cudaStream_t stream) {
    int dev = 0, limit = 0, variant = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (!choose_cluster_shape(n_sites, &a.shape, limit - 256, &variant)) {
        set_error("euler kernel: 2N=%d does not fit a cluster of %d CTAs", 2 * n_sites, MAX_CLUSTER);
        return -1;
    }
    EulerKernel fn = backward ? kEulerVariants[variant].bwd : kEulerVariants[variant].fwd;
    const int smem = smem_layout(a.shape, n_sites).total;
    SSN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    a.work_counter = counter;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.shape.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(a.shape.csize, 1, 1);
    cfg.blockDim = dim3(kEulerVariants[variant].threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    SSN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg));
    if (max_clusters < 1) { set_error("euler kernel: no resident cluster"); return -1; }
    cfg.gridDim = dim3(std::min(max_clusters, nz) * a.shape.csize, 1, 1);
    SSN_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    {
        KernelTimer kt(backward ? "ssn_euler_cluster_kernel_bwd" : "ssn_euler_cluster_kernel_fwd", stream);
        SSN_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
    }
    count_launch();
    return 0;
}

// dL/dtheta contraction (accumulates into grad): tcgen05 (TF32 x 3, TMA) by default, the FFMA kernel with SSN_K4B=ffma
static int bptt_param_grad(int nz, int nb, int n_sites, int seqlen, const float *adj, const float *traj, const float *z,
                           const WeightConst &wc, double *grad, cudaStream_t stream) {
    const int pitch = traj_pitch(n_sites);
    const char *k4b = getenv("SSN_K4B");
    if (!(k4b && !strcmp(k4b, "ffma"))) {
        const int rc = launch_bptt_param_grad_tc(nz, n_sites, (long long)seqlen * nb, pitch, adj, traj, z, wc, grad, stream);
        if (rc != 1) return rc;                                   // 1: tensor maps unavailable -> FFMA kernel
    }
    const int dim = 2 * n_sites, tiles = (dim + GT - 1) / GT;
    {
        KernelTimer kt("ssn_bptt_param_grad_kernel", stream);
        ssn_bptt_param_grad_kernel<<<nz * tiles * tiles, 256, 0, stream>>>(
            n_sites, pitch, (long long)seqlen * nb, adj, traj, z, wc, grad);
        SSN_CUDA(cudaGetLastError());
    }
    count_launch();
    return 0;
}

static void fill_common(EulerArgs &a, const ssn_solver &sv, int nz, int nb, int n_sites, const float *z,
                        const ssn_jds &jds, int seqlen, int skip, double threshold) {
    a.nz = nz; a.nb = nb; a.n_sites = n_sites;
    a.z = z; a.wc = make_weight_const(jds, n_sites);
    a.seqlen = seqlen; a.skip = skip; a.threshold = (float)threshold;
    a.pitch = traj_pitch(n_sites);
    a.io = make_io_const<float>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.eps_E = sv.dt / sv.tau_E; a.eps_I = sv.dt / sv.tau_I;
}

int launch_euler_forward(const ssn_solver &sv, int nz, int nb, int n_sites, const float *z, const ssn_jds &jds,
                         const float *ext, int ext_per_network, int seqlen, int skip_steps, double threshold,
                         float *time_avg, double *penalties, float *traj, float *gain, int *counter,
                         cudaStream_t stream) {
    SSN_CUDA(cudaMemsetAsync(penalties, 0, 2 * sizeof(double), stream));
    if (nz <= 0 || nb <= 0) return 0;
    EulerArgs a = {};
    fill_common(a, sv, nz, nb, n_sites, z, jds, seqlen, skip_steps, threshold);
    a.ext = ext; a.ext_stride_z = ext_per_network ? (long long)nb * 2 * n_sites : 0;
    a.time_avg = time_avg; a.penalties = penalties; a.traj = traj; a.gain = gain;
    return launch_euler(false, a, n_sites, nz, counter, stream);
}

int launch_euler_backward(const ssn_solver &sv, int nz, int nb, int n_sites, const float *z, const ssn_jds &jds,
                          int seqlen, int skip_steps, double threshold, const float *grad_time_avg,
                          double w_dyn, double w_rate, const float *w_dev, const float *traj, const float *gain, float *adj,
                          double *grad, float *grad_ext, int *counter, cudaStream_t stream) {
    SSN_CUDA(cudaMemsetAsync(grad, 0, 12 * sizeof(double), stream));
    if (nz <= 0 || nb <= 0) return 0;
    EulerArgs a = {};
    fill_common(a, sv, nz, nb, n_sites, z, jds, seqlen, skip_steps, threshold);
    a.g_avg = grad_time_avg; a.w_dyn = w_dyn; a.w_rate = w_rate; a.w_dev = w_dev;
    a.traj_in = traj; a.gain_in = gain; a.adj = adj; a.grad_ext = grad_ext;
    int rc = launch_euler(true, a, n_sites, nz, counter, stream);
    if (rc) return rc;
    return bptt_param_grad(nz, nb, n_sites, seqlen, adj, traj, z, a.wc, grad, stream);
}

int launch_bptt_param_grad(int nz, int nb, int n_sites, int seqlen, const float *adj, const float *traj, const float *z,
                           const ssn_jds &jds, double *grad, cudaStream_t stream) {
    SSN_CUDA(cudaMemsetAsync(grad, 0, 12 * sizeof(double), stream));
    if (nz <= 0) return 0;
    return bptt_param_grad(nz, nb, n_sites, seqlen, adj, traj, z, make_weight_const(jds, n_sites), grad, stream);
}

}  // namespace ssn
