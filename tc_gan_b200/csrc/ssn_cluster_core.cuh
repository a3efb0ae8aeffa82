// Building blocks shared by the cluster-resident SSN kernels (sm_100a).
//
// One thread-block cluster owns one network.  CTA `rank` of a cluster of `csize`
// CTAs keeps rows [rank*rpc, rank*rpc+rpc) of the 2N x 2N matrix (W, or W^T for
// the adjoint kernels) in its shared memory for the whole life of the network,
// and the full state panel X[2N][8] (8 stimuli wide) in a double-buffered,
// bank-conflict-free float4 layout that every CTA of the cluster updates through
// distributed shared memory once per sweep.
//
// Thread mapping inside a CTA: thread = (row group g, k-lane kl), lane = g_in_warp*KL + kl.
// A thread accumulates a TI x 8 register tile over the columns j = 4*(s*KL+kl)+q,
// q = 0..3, s = 0..kpad/(4 KL)-1:  one LDS.128 of W per row per step (the 8 lanes
// of a quarter warp read 128 contiguous bytes), and two broadcast LDS.128 of X
// per column (all row groups of a warp read the same address).  The KL partial
// tiles are then reduce-scattered with warp shuffles so that lane kl ends up
// with the TI finished sums of stimulus kl.
#pragma once
#include <cooperative_groups.h>
#include "ssn_common.cuh"

namespace ssn {
namespace cg = cooperative_groups;

constexpr int TB = 8;          // stimuli per panel
constexpr int MAX_CLUSTER = 8;

// Panel layout: plane h (stimuli 4h..4h+3) of buffer `buf` is P float4, column j
// lives at float4 index 5*(j>>2) + (j&3): the +1 skew every four columns makes
// the eight k-lanes of a quarter warp (columns 4 kl + q) hit eight distinct
// 16-byte bank groups.
__host__ __device__ inline int panel_P(int kpad) { return 5 * (kpad / 4); }
__device__ __forceinline__ int panel_col(int j) { return 5 * (j >> 2) + (j & 3); }
// float index of (buf, column j, stimulus b)
__device__ __forceinline__ int panel_index(int P, int buf, int j, int b) {
    return (((buf * 2 + (b >> 2)) * P + panel_col(j)) << 2) + (b & 3);
}

struct ClusterShape {
    int dim, kpad, csize, rpc;        // kpad = round_up(dim, 32); rpc = rows per CTA
};

// Shared-memory carve-up (bytes, all 16-byte aligned).
struct SmemLayout {
    int w_off, x_off, ext_off, gtab_off, misc_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(const ClusterShape &s, int n_sites) {
    SmemLayout L;
    int o = 0;
    L.w_off = o;    o += s.rpc * s.kpad * 4;
    L.x_off = o;    o += 2 * 2 * panel_P(s.kpad) * 16;
    L.ext_off = o;  o += ((s.rpc * TB * 4 + 15) / 16) * 16;
    L.gtab_off = o; o += ((4 * n_sites * 4 + 15) / 16) * 16;
    L.misc_off = o; o += 1024;
    L.total = o;
    return L;
}

// misc block: flags[2][MAX_CLUSTER] (uint32), myflags, next_net
struct Misc {
    unsigned flags[2][MAX_CLUSTER];
    unsigned myflags;
    int next_net;
    float scale[2][MAX_CLUSTER][TB];   // per-rank per-stimulus scratch (adjoint norms)
};
static_assert(sizeof(Misc) <= 1024, "misc block");

// ---- W slice loaders -------------------------------------------------------------
// Rows [row_base, row_base+rows_here) of W (or of W^T when TRANSPOSED) into Wsm[r*kpad + j].
// src is W itself (w_kind = SSN_W_DENSE) or z (SSN_W_FROM_Z), both [dim][dim] row-major.
template <bool TRANSPOSED>
__device__ __forceinline__ void load_matrix_slice(float *Wsm, const float *__restrict__ src, int w_kind,
                                                  const WeightConst &wc, const float *gtab,
                                                  int n_sites, int dim, int kpad,
                                                  int row_base, int rows_here, int tid, int nthreads) {
    const int total = rows_here * kpad;
    if (!TRANSPOSED) {
        // consecutive threads walk along a row of W: coalesced global reads, conflict-free stores
#pragma unroll 16
        for (int idx = tid; idx < total; idx += nthreads) {
            const int r = idx / kpad, j = idx - r * kpad;
            const int i = row_base + r;
            float v = 0.f;
            if (j < dim) {
                v = __ldg(src + (size_t)i * dim + j);
                if (w_kind == SSN_W_FROM_Z) v = weight_from_z(wc, gtab, n_sites, i, j, v);
            }
            Wsm[idx] = v;
        }
    } else {
        // slice row r holds column (row_base + r) of W: Wsm[r][j] = W[j][row_base + r].
        // Consecutive threads take consecutive r, so global reads still run along a row of W.
#pragma unroll 16
        for (int idx = tid; idx < total; idx += nthreads) {
            const int j = idx / rows_here, r = idx - j * rows_here;
            float v = 0.f;
            if (j < dim) {
                v = __ldg(src + (size_t)j * dim + row_base + r);
                if (w_kind == SSN_W_FROM_Z) v = weight_from_z(wc, gtab, n_sites, j, row_base + r, v);
            }
            Wsm[r * kpad + j] = v;
        }
    }
}

// ---- the skinny contraction --------------------------------------------------------
// acc[t][b] = sum_j Wsm[row(t)][j] * X[buf][j][b]  over this thread's columns.
// The accumulators are packed stimulus pairs (b, b+1): one fma.rn.f32x2 per pair with the broadcast W element, its
// x operand being the register pair the panel's LDS.128 delivered -- half the issue slots of scalar FFMAs and no
// register-bank conflicts (the scalar version measured 54 % of its FFMAs with a 2-cycle conflict).  Measured on the
// BPTT step at configs[2]: scalar FFMA + W rows of the next column step requested early 53.2 ms, scalar 52.5,
// packed + early W 50.1, packed 47.8 (kept) -- the loop is bound by the 15 LDS.128 per 112 packed FMAs (60
// wavefronts per warp and column step), not by issue slots, and the prefetch only costs registers.
__device__ __forceinline__ unsigned long long bcast2(float x) {
    unsigned long long v;
    asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(x));
    return v;
}
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
template <int TI, int KL>
__device__ __forceinline__ void contract_panel(float (&acc)[TI][TB], const float *Wsm, const float4 *X4,
                                               int P, int kpad, int buf, const int (&wrow)[TI], int kl) {
    const int nsteps = kpad / (4 * KL);
    const int kp4 = kpad / 4;
    const float4 *wp[TI];
#pragma unroll
    for (int t = 0; t < TI; ++t) wp[t] = reinterpret_cast<const float4 *>(Wsm) + wrow[t] * kp4 + kl;
    unsigned long long a2[TI][TB / 2];
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int b = 0; b < TB / 2; ++b) a2[t][b] = 0ull;
    const ulonglong2 *Xa = reinterpret_cast<const ulonglong2 *>(X4 + (buf * 2 + 0) * P + 5 * kl);
    const ulonglong2 *Xb = Xa + P;
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        float4 w[TI];
#pragma unroll
        for (int t = 0; t < TI; ++t) w[t] = wp[t][s * KL];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const ulonglong2 xa = Xa[s * 5 * KL + q];
            const ulonglong2 xb = Xb[s * 5 * KL + q];
#pragma unroll
            for (int t = 0; t < TI; ++t) {
                const float wq = q == 0 ? w[t].x : q == 1 ? w[t].y : q == 2 ? w[t].z : w[t].w;
                const unsigned long long ww = bcast2(wq);
                fma2(a2[t][0], ww, xa.x);
                fma2(a2[t][1], ww, xa.y);
                fma2(a2[t][2], ww, xb.x);
                fma2(a2[t][3], ww, xb.y);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int b = 0; b < TB / 2; ++b)
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[t][2 * b]), "=f"(acc[t][2 * b + 1]) : "l"(a2[t][b]));
}

// ---- ownership after the reduce-scatter ------------------------------------------------
// The KL k-lanes of a row group hold partial sums of a TI x 8 tile.  After the
// reduce-scatter, lane kl owns stimulus kl / (KL/8) and the rows
// [sub*TO, sub*TO+TO) of the group, sub = kl % (KL/8), TO = ceil(TI / (KL/8)).
template <int TI, int KL>
struct Owner {
    static constexpr int SPLIT = KL / TB;                  // lanes sharing one stimulus
    static constexpr int TO = (TI + SPLIT - 1) / SPLIT;    // rows owned per lane
    __device__ static __forceinline__ int stim(int kl) { return kl / SPLIT; }
    __device__ static __forceinline__ int first_row(int kl) { return (kl % SPLIT) * TO; }
};

// Reduce-scatter of the KL partial tiles (KL = 8 or 16): on return out[u] is the full
// sum for stimulus Owner::stim(kl), row Owner::first_row(kl) + u of the group.
template <int TI, int KL>
__device__ __forceinline__ void reduce_scatter(const float (&acc)[TI][TB],
                                               float (&out)[Owner<TI, KL>::TO], int kl) {
    static_assert(KL == 8 || KL == 16, "k-lanes");
    const unsigned full = 0xffffffffu;
    constexpr int S = KL / TB;                 // 1 or 2
    float v4[TI][4], v2[TI][2], v1[TI];
    const bool u4 = kl & (4 * S), u2 = kl & (2 * S), u1 = kl & S;
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float send = u4 ? acc[t][c] : acc[t][4 + c];
            const float keep = u4 ? acc[t][4 + c] : acc[t][c];
            v4[t][c] = keep + __shfl_xor_sync(full, send, 4 * S);
        }
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float send = u2 ? v4[t][c] : v4[t][2 + c];
            const float keep = u2 ? v4[t][2 + c] : v4[t][c];
            v2[t][c] = keep + __shfl_xor_sync(full, send, 2 * S);
        }
#pragma unroll
    for (int t = 0; t < TI; ++t) {
        const float send = u1 ? v2[t][0] : v2[t][1];
        const float keep = u1 ? v2[t][1] : v2[t][0];
        v1[t] = keep + __shfl_xor_sync(full, send, S);
    }
    if (S == 1) {
#pragma unroll
        for (int u = 0; u < Owner<TI, KL>::TO; ++u) out[u] = v1[u < TI ? u : 0];
    } else {
        // two lanes hold partial sums of the same stimulus: split the rows between them
        constexpr int TO = Owner<TI, KL>::TO;
        const bool hi = kl & 1;
#pragma unroll
        for (int u = 0; u < TO; ++u) {
            const float mine = hi ? (TO + u < TI ? v1[TO + u < TI ? TO + u : 0] : 0.f) : v1[u];
            const float theirs = hi ? v1[u] : (TO + u < TI ? v1[TO + u < TI ? TO + u : 0] : 0.f);
            out[u] = mine + __shfl_xor_sync(full, theirs, 1);
        }
    }
}

// OR of a per-thread predicate over the lanes that own the same stimulus, as a bit
// mask indexed by stimulus (lane l owns stimulus (l % KL) / (KL/8)).
template <int KL>
__device__ __forceinline__ unsigned stim_mask(bool pred) {
    unsigned m = __ballot_sync(0xffffffffu, pred);
    if (KL == 8) {
        m |= m >> 16;
        m |= m >> 8;
        return m & 0xffu;
    }
    m |= m >> 16;                 // fold the two row groups of the warp
    m &= 0xffffu;
    m |= m >> 1;                  // fold the two lanes of a stimulus: bits 0,2,4,..
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < TB; ++b) r |= ((m >> (2 * b)) & 1u) << b;
    return r;
}

// ---- distributed shared memory helpers (32-bit shared::cluster addresses) ----------------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ unsigned map_to_rank(unsigned smem_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(unsigned addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(unsigned addr, unsigned v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// fast float transfer function: MUFU-based pow / tanh (relative error ~4e-7), used where
// the float64 reference expansion does not apply
__device__ __forceinline__ float io_eval_fast(const IoConst<float> &c, float v) {
    if (v <= 0.f) return 0.f;
    if (c.io_type == SSN_IO_POWER || v <= c.v0) return c.k * exp2f(c.n * __log2f(v));
    if (c.io_type == SSN_IO_LINEAR) return fmaf(c.lin_slope, v - c.v0, c.r_soft);
    const float e = __expf(2.f * c.tanh_scale * (v - c.v0));        // tanh(x) = 1 - 2/(e^{2x}+1)
    return fmaf(c.span, 1.f - __fdividef(2.f, e + 1.f), c.r_soft);
}


// Power-law branch f = k v^n and f' = n k v^(n-1) for v > 0 from the MUFU lg2 / ex2 units: the integer part of n is
// taken by multiplications and only v^frac(n) goes through 2^(frac log2 v), so the approximate exponent stays small
// (|frac(n) log2 v| < 2 for n = 2.2, v < 1000) and the relative error is ~2e-7 instead of ~1e-6 for 2^(n log2 v);
// no division, no branches.  (powf: 1 ulp, but ~60 instructions with slow paths, once per neuron, stimulus and step.)
__device__ __forceinline__ void io_power_fast(const IoConst<float> &c, float v, float &f, float &df) {
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(c.n_frac * __log2f(v)));
    // v^(n_int - 1) from the bits of the exponent (n_int <= 8): straight-line code, no loop
    const int e = c.n_int - 1;
    const float v2 = v * v, v4 = v2 * v2;
    float vm = (e & 1) ? v : 1.f;
    vm = (e & 2) ? vm * v2 : vm;
    vm = (e & 4) ? vm * v4 : vm;
    if (c.n_int >= 1) {
        df = c.nk * vm * p;
        f = c.k * (vm * v) * p;
    } else {
        f = c.k * p;
        df = __fdividef(c.n * f, v);
    }
}

}  // namespace ssn
