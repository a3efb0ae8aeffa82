// K1: warp-specialised fixed-point solve with the weight matrix resident in registers.
//
// One thread-block cluster owns one network (persistent clusters pull networks from an atomic counter).  CTA
// `rank` owns rows [rank * rpc, +rpc) of W; inside a CTA
//
//   * CW CONTRACTION warps hold the W tile (TI rows x NC columns per thread: warp w owns rows w*TI.., lane l the
//     columns l, l+32, ...; rows packed in pairs for FFMA2) and do nothing but  wait panel -> FFMA2 contraction ->
//     32-lane reduce-scatter -> hand one dv per lane to their update warp through shared memory;
//   * UW UPDATE warps (warp u serves the G = CW / UW contraction warps u, u+UW, ..: the ones on its own scheduler; lane = 4 * row + stimulus
//     slot owns one output per served warp and stream) own the float64 state, evaluate f from the tables, apply
//     the Euler step and the stopping tests, and publish the new r - r_ref (and the warp's flag word) to every CTA
//     of the cluster with st.async remote stores that complete bytes on the destination's mbarrier;
//   * the stimuli of a network run as TWO independent streams of half-panels (4 stimuli), each stream with its own
//     double-buffered panel and mbarriers.  While the update warps and the cluster exchange finish sweep k of one
//     stream, the contraction warps are already in sweep k of the other, so the float64 update, the publish and
//     the DSMEM latency run under FMA work instead of after it.  A stream whose half-panel has converged picks up
//     the next half-panel of the network (nb > 8), so both streams stay busy until the network runs out of
//     stimuli; only then does the last stream run alone.
//
// Shapes.  The kernel is generic in <CW, UW, TI>.  Production shape <8, 8, 7>: 56 rows per CTA, a cluster of 8
// at 2N = 402 (15 resident clusters = 120 SMs on a B200), one update warp per contraction warp.  The wider
// <12, 4, 6> (72 rows per CTA: clusters of 6, 22 resident = 132 SMs, three contraction warps per scheduler) is
// compiled only with -DSSN_WS_SHAPE_B: measured 96 k solves/s against 111 k -- its four update warps carry three
// outputs per lane and become the bottleneck of every sweep.
//
// Numerics -- reference-point iteration.  Per stimulus the kernel iterates on r - r_ref:
// v = v_ref + W fl32(r - r_ref) (FP32 FFMA2), f(v) in float64 from cubic Taylor tables in shared memory,
// r <- r + eps (f - r) in float64.  r_ref starts as the initial state (v_ref = I exactly for r_init = 0) and is
// refreshed -- one exact W r_ref per stimulus: float64 accumulation of exact fp32 x fp32 products of hi/lo split
// states -- every time max|dr| has shrunk 64-fold, so the FP32 contraction error stays ~1e-7 RELATIVE TO THE
// REMAINING DISTANCE to the fixed point and the sweep at which |r_new - r_old| < atol first holds is the one of
// the float64 reference solver.
//
// Register budget: the CTA starts with 128 registers per thread (512 threads); the update warp group(s) release
// down to R_U and the contraction warp groups grow to R_C with setmaxnreg (32 CW R_C + 32 UW R_U <= 65536; the
// roles are whole warp groups of 4, as setmaxnreg requires).  (This file must NOT be compiled with -rdc: ptxas
// ignores setmaxnreg in relocatable device code.)
//
// Stopping rule and error codes follow tc_gan/ext/ssnode.c:84-102: converged is tested before the hard bound,
// the tanh transfer function has no hard-bound exit.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "ssn_ws_common.cuh"
#include "ssn_launch.h"

#ifndef SSN_WS_PROFILE
#define SSN_WS_PROFILE 0
#endif
// tuning switches (kept for the measurements quoted in DESIGN.md)
#ifndef SSN_WS_SPEC
#define SSN_WS_SPEC 1        // contract speculatively, evaluate the flag word afterwards
#endif
#ifndef SSN_WS_PF
#define SSN_WS_PF 2          // panel columns loaded this many columns ahead of use
#endif
#ifndef SSN_WS_DIAG
#define SSN_WS_DIAG 0        // timing diagnostics (WRONG results): 1 no FMAs, 2 no reduce-scatter, 4 trivial f, 8 no remote publish of values
#endif
#ifndef SSN_WS_RED
#define SSN_WS_RED 0         // 0: 32-lane shuffle reduce-scatter of the 7 x 4 tile (default); 1: transpose through shared
                             // memory -- half the instructions (67 against 124), measured 0.8 % SLOWER (74.16 against
                             // 73.59 ms at configs[1]): the sweep is bound by the latency of its chain, not by issue slots
#endif
constexpr int WS_RED_PITCH = 36;                        // floats per output row of the transpose (32 lanes + skew)

namespace ssn {

constexpr int WS_THREADS = 512;                         // 16 warps in every shape
constexpr int WS_BAR_REFRESH = 1;                       // named barrier used by refresh events (all threads)
constexpr int WS_MAX_CW = 12;

// Compile-time shape of a CTA.
template <int CW_, int UW_, int TI_, int RC_, int PF_ = SSN_WS_PF>
struct WsShape {
    static constexpr int PF = PF_;                      // panel columns loaded this many columns ahead of use
    static constexpr int CW = CW_, UW = UW_, TI = TI_;
    static constexpr int G = CW / UW;                   // contraction warps served by one update warp
    static constexpr int NP = TI / 2;                   // packed row pairs
    static constexpr bool ODD = TI & 1;                 // one single row left over
    static constexpr int ROWS = CW * TI;                // row slots per CTA
    static constexpr int REG_C = RC_;
    static constexpr int REG_U = ((65536 - 32 * CW * RC_) / (32 * UW)) / 8 * 8 > 128 ? 128 : ((65536 - 32 * CW * RC_) / (32 * UW)) / 8 * 8;
    static_assert(CW + UW == 16 && CW % 4 == 0 && UW % 4 == 0, "roles are whole warp groups of a 16-warp CTA");
    static_assert(CW % UW == 0, "an update warp serves a whole number of contraction warps");
    static_assert(4 * TI <= 28, "lane = 4 * row + stimulus, lane 28 carries the flag word");
    static_assert(MAX_CLUSTER * UW <= 64, "one or two flag words per lane");
    static_assert(CW <= WS_MAX_CW, "misc block");
    static_assert(REG_U >= 40, "update warps need registers too");
};
// State panel of a stream (one buffer): kpad 16-byte slots, slot j = the float4 (four stimuli) of global row j
// (rows dim..kpad-1 stay zero), then MAX_CLUSTER * UW flag words (one per update warp of every CTA).  A lane's
// column c is at  lane * 16 + c * 512  from the buffer: one base register and immediate offsets, and every
// LDS.128 of a warp reads 512 contiguous bytes.
__host__ __device__ constexpr int ws_buf_bytes(int kpad) { return kpad * 16 + 256; }

struct WsMisc {
    unsigned long long full[2][2];          // [half][buffer]: panel of the next sweep complete
    unsigned long long xfull[2];            // [half]: hi/lo columns of a refresh event complete
    unsigned long long dvfull[2][WS_MAX_CW];// [half][contraction warp]: dv handed over
    double tlevel[8];                       // refresh ladder thresholds by level, 0 = exhausted
    unsigned pdelta[MAX_CLUSTER];
    int next_net;
};
static_assert(sizeof(WsMisc) <= 512, "misc block");

struct WsSmem {
    int x_off, xe_off, tab_off, gtab_off, dv_off, ex_off, ref_off, red_off, misc_off, total;
};
template <class SH>
__host__ __device__ inline WsSmem ws_smem_layout(int kpad, int n_sites, int tab_bytes) {
    WsSmem L;
    int o = 0;
    L.x_off = o;    o += 4 * ws_buf_bytes(kpad);            // [half][buffer]
    L.xe_off = o;   o += 2 * 2 * 4 * kpad * 4;              // [half][hi, lo][4 stimuli][kpad]
    L.tab_off = o;  o += tab_bytes;                         // Taylor tables of f
    L.gtab_off = o; o += ((4 * n_sites * 4 + 15) / 16) * 16;
    L.dv_off = o;   o += 2 * SH::CW * 32 * 4;               // [half][contraction warp][lane] float
    L.ex_off = o;   o += 2 * SH::CW * 32 * 8;               // [half][contraction warp][lane] double
    L.ref_off = o;  o += 2 * 2 * SH::CW * 32 * 8;           // [half][r_ref, v_ref][contraction warp][lane] double
    L.red_off = o;  o += SSN_WS_RED ? SH::CW * 4 * SH::TI * WS_RED_PITCH * 4 : 0;   // [contraction warp][4 TI outputs][pitch]
    L.misc_off = o; o += 512;
    L.total = o;
    return L;
}

template <int R> __device__ __forceinline__ void reg_grow() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_release() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ void bar_sync_all(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(WS_THREADS) : "memory"); }

template <int H> using HalfC = std::integral_constant<int, H>;
// bit (in the done / force masks and the flag fields) of slot b (0..3) of stream h, and the stream's bit mask
__host__ __device__ constexpr int ws_stim(int h, int b) { return 4 * h + b; }
__host__ __device__ constexpr unsigned ws_mask(int h) { return 0xfu << (4 * h); }

// OR of the flag words of all update warps of all CTAs in a panel buffer (absent CTAs stay zero)
template <class SH>
__device__ __forceinline__ unsigned ws_flag_word(const unsigned char *buf_base, int kpad, int lane) {
    if (SH::UW == 8) {                  // 64 words: two per lane
        const uint2 f = *reinterpret_cast<const uint2 *>(buf_base + kpad * 16 + lane * 8);
        return f.x | f.y;
    }
    // UW == 4: 32 words, one per lane
    return *reinterpret_cast<const unsigned *>(buf_base + kpad * 16 + lane * 4);
}

template <int NC, class SH>
__global__ void __launch_bounds__(WS_THREADS, 1) ssn_fp_ws_kernel(const RwArgs a) {
    constexpr int CW = SH::CW, UW = SH::UW, TI = SH::TI, NP = SH::NP, G = SH::G;
    constexpr bool ODD = SH::ODD;
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int csize = a.csize, dim = a.dim, kpad = a.kpad, rpc = a.rpc, N = a.n_sites;
    constexpr int BUF_BYTES = ws_buf_bytes(32 * NC);
    const WsSmem L = ws_smem_layout<SH>(kpad, N, rw_table_bytes(a.tab_nodes, a.tab2_nodes));
    float *xe = reinterpret_cast<float *>(smem + L.xe_off);
    double *tab = reinterpret_cast<double *>(smem + L.tab_off);
    float *gtab = reinterpret_cast<float *>(smem + L.gtab_off);
    float *dvbuf = reinterpret_cast<float *>(smem + L.dv_off);
    double *exbuf = reinterpret_cast<double *>(smem + L.ex_off);
    double *refbuf = reinterpret_cast<double *>(smem + L.ref_off);
    WsMisc *misc = reinterpret_cast<WsMisc *>(smem + L.misc_off);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = rank * rpc;
    const int rows_here = max(0, min(rpc, dim - row_base));
    const unsigned x_local = smem_u32(smem + L.x_off), xe_local = smem_u32(xe);
    const int sid = lane & 3;                                       // slot (of either stream) whose status this lane tracks

    // ---- one-time setup (all warps) ----
    if (tid < MAX_CLUSTER) misc->pdelta[tid] = map_to_rank(x_local, (unsigned)(tid < csize ? tid : 0)) - x_local;
    if (tid >= 32 && tid < 40) {
        const int l = tid - 32;
        double t = a.t_first;
        for (int q = 0; q < l; ++q) t *= (1.0 / 64.0);
        misc->tlevel[l] = (l == 7 || t <= a.atol) ? 0.0 : t;
    }
    if (tid == 0) {
        for (int h = 0; h < 2; ++h) {
            mbar_init(smem_u32(&misc->full[h][0]), 1 + UW);        // the arming thread + one release-arrive per update warp
            mbar_init(smem_u32(&misc->full[h][1]), 1 + UW);
            mbar_init(smem_u32(&misc->xfull[h]), 1);
            for (int w = 0; w < CW; ++w) mbar_init(smem_u32(&misc->dvfull[h][w]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 4 * BUF_BYTES / 4; i += WS_THREADS) reinterpret_cast<float *>(smem + L.x_off)[i] = 0.f;
    for (int i = tid; i < 2 * 2 * 4 * kpad; i += WS_THREADS) xe[i] = 0.f;
    if (a.w_kind == SSN_W_FROM_Z) build_profile_table(a.wc, N, gtab, tid, WS_THREADS);
    build_io_tables(a, tab, tid, WS_THREADS);
    cluster.sync();

    const volatile unsigned *pdelta = misc->pdelta;
    // bytes arriving from the peers per panel: 16 per row they own + their update warps' flag words
    const unsigned tx_bytes = (unsigned)((dim - rows_here) * 16 + (csize - 1) * UW * 4);
    const int n_hp = (a.nb + 3) / 4;                        // half-panels (4 stimuli) of a network: the units the streams pull
    auto full_bar = [&](int h, int b) { return smem_u32(&misc->full[h][b]); };
    auto panel = [&](int h, int b) { return smem + L.x_off + (h * 2 + b) * BUF_BYTES; };

    // Per-stream sweep bookkeeping, computed identically by both roles from the cluster-uniform flag word F.
    // Returns false when the stream's half-panel has finished.  `it` is the half-panel's sweep counter.
    //   done / force: 8-bit masks, bits 4h..4h+3 = the four stimulus slots of stream h (hm selects them).
    // Both roles walk the streams in the same order (0, 1, 0, 1, ... over the live ones), so every decision taken
    // here -- including which half-panel a finished stream picks up next -- is the same in all warps of the cluster.
    auto advance = [&](unsigned F, unsigned hm, int it, unsigned &done, unsigned &force, unsigned &req,
                       unsigned &natural, unsigned &conv_now, unsigned &hard_now) -> bool {
        conv_now = 0u; hard_now = 0u;
        if (it > 1) {
            const unsigned moving_all = F & 0xffu, above_all = (F >> 8) & 0xffu;
            conv_now = ~moving_all & ~done & hm;                                   // ssnode.c:84-96 first ...
            hard_now = a.check_hard ? (above_all & ~done & ~conv_now & hm) : 0u;   // ... then :98-102
            done |= conv_now | hard_now;
        }
        if ((done & hm) == hm || it > a.max_iter) return false;
        natural = ~(F >> 16) & ~done & hm;
        req = natural | (force & ~done & hm);
        force &= ~req;
        return true;
    };

    // slots of a stream working on half-panel hp that hold no stimulus (bits 0..3)
    auto empty_slots = [&](int hp) -> unsigned {
        const int n = a.nb - 4 * hp;
        return n >= 4 ? 0u : (n <= 0 ? 0xfu : (0xfu << n) & 0xfu);
    };

    if (warp < CW) {
        // =====================================================================================
        // contraction warps
        // =====================================================================================
        reg_grow<SH::REG_C>();
        const int cwarp = warp, ctid = tid;
        const int row0 = cwarp * TI;
        unsigned ph = 0u, xph = 0u;                                 // parity bits: full[h][b] -> bit 2h+b, xfull[h] -> bit h
#if SSN_WS_PROFILE
        long long tc[6] = {0, 0, 0, 0, 0, 0};
#endif
        for (;;) {
            if (rank == 0 && ctid == 0) {
                const int n = atomicAdd(a.work_counter, 1);
                for (int p = 0; p < csize; ++p) st_cluster_u32(map_to_rank(smem_u32(&misc->next_net), p), (unsigned)n);
            }
            cluster.sync();
            const int net = misc->next_net;
            if (net >= a.nz) break;

            // ---- W tile -> registers, one row pair at a time (a whole-tile staging array would not fit beside
            //      the packed tile and spilled in the first version) ----
            unsigned long long wp[NP > 0 ? NP : 1][NC];
            float ws[NC];
            {
                // `lane_o` is the lane behind an opaque zero defined inside the network loop: without it ptxas hoists
                // the per-column table addresses and predicates of this once-per-network prologue out of the
                // persistent loop and keeps ~25 registers live through every sweep
                int lane_o;
                asm volatile("mov.u32 %0, %1;" : "=r"(lane_o) : "r"(lane));
                const float *src = a.w + (size_t)net * dim * dim + (size_t)(row_base + row0) * dim + lane_o;
                auto load_row = [&](int t, float (&zv)[NC]) {
                    const bool row_ok = row0 + t < rows_here;
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        zv[c] = (row_ok && c * 32 + lane_o < dim) ? __ldg(src + (size_t)t * dim + c * 32) : 0.f;
                };
                auto make_w = [&](int t, float (&zv)[NC]) {
                    if (a.w_kind != SSN_W_FROM_Z) return;
                    const int i = row_base + row0 + t;
                    const int ah = i >= N, ii = i - ah * N;
                    const bool row_ok = row0 + t < rows_here;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int j = c * 32 + lane_o;
                        const int bh = j >= N, ab = ah * 2 + bh;
                        int d = ii - (j - bh * N);
                        d = d < 0 ? -d : d;
                        zv[c] = (row_ok && j < dim) ? gtab[ab * N + min(d, N - 1)] * fmaf(a.wc.sD[ab], zv[c], a.wc.sJ[ab]) : 0.f;
                    }
                };
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    float z0[NC], z1[NC];
                    load_row(2 * q, z0);
                    load_row(2 * q + 1, z1);
                    make_w(2 * q, z0);
                    make_w(2 * q + 1, z1);
#pragma unroll
                    for (int c = 0; c < NC; ++c) wp[q][c] = pack2(z0[c], z1[c]);
                }
                if (ODD) {
                    load_row(TI - 1, ws);
                    make_w(TI - 1, ws);
                } else {
#pragma unroll
                    for (int c = 0; c < NC; ++c) ws[c] = 0.f;
                }
            }

            {
                // stream h starts on half-panel h; a stream that finishes picks up the next one of this network
                int next_hp = min(2, n_hp);
                unsigned done = empty_slots(0) | (empty_slots(1) << 4);
                unsigned force = a.r_init ? (~done & 0xffu) : 0u;
                unsigned alive = n_hp > 1 ? 3u : 1u;
                unsigned bufbits = 0u;
                int it0 = 1, it1 = 1;
                // the stream's half-panel is finished: take the next one (after a cluster barrier: nobody may publish
                // the new panel while a slower CTA still looks at the flags of the old one) or retire the stream
                auto c_finish = [&](int h) {
                    if (next_hp < n_hp) {
                        cluster.sync();
                        const unsigned hm = ws_mask(h);
                        done = (done & ~hm) | (empty_slots(next_hp) << (4 * h));
                        force = (force & ~hm) | (a.r_init ? (~done & hm) : 0u);
                        ++next_hp;
                        (h ? it1 : it0) = 1;
                        bufbits ^= 1u << h;                 // the new initial panel arrives in the other buffer
                    } else {
                        alive &= ~(1u << h);
                    }
                };

                auto cstep = [&](auto hc) {
                    constexpr int h = decltype(hc)::value;
                    constexpr unsigned hm = ws_mask(h);
                    const int buf = (bufbits >> h) & 1u;
                    int &it = h ? it1 : it0;
#if SSN_WS_PROFILE
                    long long c0 = clock64();
#endif
                    mbar_wait(full_bar(h, buf), (ph >> (2 * h + buf)) & 1u);
                    ph ^= 1u << (2 * h + buf);
#if SSN_WS_PROFILE
                    long long c1 = clock64(); tc[0] += c1 - c0;
#endif
                    unsigned char *xb = panel(h, buf);
                    // the flag word is loaded now and looked at after the contraction: finishing and refresh
                    // events are rare, so the contraction runs speculatively under the latency of the flag logic
                    const unsigned fword = ws_flag_word<SH>(xb, kpad, lane);
                    unsigned long long ap[NP > 0 ? NP : 1][4];
                    float as[4];
                    // ---- contraction of the half: dv = W * fl32(r - r_ref), four stimuli per LDS.128,
                    //      panel columns fetched SSN_WS_PF columns ahead of their use ----
                    auto contract = [&]() {
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q = 0; q < NP; ++q) ap[q][b] = 0ull;
                            as[b] = 0.f;
                        }
                        const float4 *Xq = reinterpret_cast<const float4 *>(xb) + lane;     // column c: Xq[32 * c]
                        constexpr int PF = SH::PF;
                        float4 xq[PF + 1];
#pragma unroll
                        for (int c = 0; c < PF && c < NC; ++c) xq[c] = Xq[32 * c];
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            if (c + PF < NC) xq[(c + PF) % (PF + 1)] = Xq[32 * (c + PF)];
                            const float4 x4 = xq[c % (PF + 1)];
                            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
#if SSN_WS_DIAG & 1
                                as[b] += xv[b];
#elif SSN_WS_DIAG & 16
                                if (ODD) as[b] = fmaf(ws[c], xv[b], as[b]);       // 1/7 of the FMA work, sane magnitudes
#else
#pragma unroll
                                for (int q = 0; q < NP; ++q) ffma2(ap[q][b], wp[q][c], xv[b]);
                                if (ODD) as[b] = fmaf(ws[c], xv[b], as[b]);
#endif
                            }
                        }
                    };
                    auto refresh = [&](unsigned req) {
                            // ---- reference-point refresh: exact W * r for the requested stimuli of this half ----
                            mbar_wait(smem_u32(&misc->xfull[h]), (xph >> h) & 1u);
                            xph ^= 1u << h;
                            for (int b = 0; b < 4; ++b) {
                                const int s = ws_stim(h, b);
                                if (!((req >> s) & 1u)) continue;
                                const int xs = __popc(req & hm & ((1u << s) - 1u));
                                const float *xh = xe + ((h * 2 + 0) * 4 + xs) * kpad, *xl = xe + ((h * 2 + 1) * 4 + xs) * kpad;
                                double accd[TI];
                                float accf[TI];
#pragma unroll
                                for (int t = 0; t < TI; ++t) { accd[t] = 0.0; accf[t] = 0.f; }
#pragma unroll
                                for (int c = 0; c < NC; ++c) {
                                    const double hv = (double)xh[c * 32 + lane];
                                    const float lv = xl[c * 32 + lane];
                                    float wv[TI];
#pragma unroll
                                    for (int q = 0; q < NP; ++q) unpack2(wp[q][c], wv[2 * q], wv[2 * q + 1]);
                                    if (ODD) wv[TI - 1] = ws[c];
#pragma unroll
                                    for (int t = 0; t < TI; ++t) {
                                        accd[t] = fma((double)wv[t], hv, accd[t]);          // exact products, fp64 sum
                                        accf[t] = fmaf(wv[t], lv, accf[t]);
                                    }
                                }
                                double mine = 0.0;
#pragma unroll
                                for (int t = 0; t < TI; ++t) {
                                    double v = accd[t] + (double)accf[t];
#pragma unroll
                                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                    mine = (t == (lane >> 2)) ? v : mine;
                                }
                                if ((lane & 3) == b) exbuf[(h * CW + cwarp) * 32 + lane] = mine;
                                // r - r_ref is now zero for this stimulus on every row of every CTA
                                float *col = reinterpret_cast<float *>(xb) + b;
                                for (int q = ctid; q < dim; q += 32 * CW) col[4 * q] = 0.f;
                            }
                            bar_sync_all(WS_BAR_REFRESH);
                    };
                    unsigned req, natural, conv_now, hard_now;
#if SSN_WS_SPEC
                    // finishing and refresh events are rare: contract first, look at the flags afterwards
                    contract();
                    {
                        const unsigned F = __reduce_or_sync(0xffffffffu, fword);
                        if (!advance(F, hm, it, done, force, req, natural, conv_now, hard_now)) { c_finish(h); return; }
                    }
                    if (req) { refresh(req); contract(); }
#else
                    {
                        const unsigned F = __reduce_or_sync(0xffffffffu, fword);
                        if (!advance(F, hm, it, done, force, req, natural, conv_now, hard_now)) { c_finish(h); return; }
                    }
                    if (req) refresh(req);
                    contract();
#endif
#if SSN_WS_PROFILE
                    long long c3 = clock64(); tc[2] += c3 - c1;
#endif
#if SSN_WS_DIAG & 2
                    float out = as[lane & 3];
                    {
                        float lo, hi;
                        unpack2(ap[0][lane & 3], lo, hi);
                        out += lo * 1e-30f;
                    }
#elif SSN_WS_RED
                    // ---- transpose through shared memory: every lane leaves its 4 TI partial sums in column `lane` of
                    //      the warp's [4 TI][36] slab (conflict-free: bank = lane), then lane 4 * row + stimulus reads
                    //      the 32 partials of its output as 8 LDS.128 (rows 36 floats apart: a quarter warp covers all
                    //      32 banks) and adds them up.  28 STS + 8 LDS.128 + 31 FADD against 31 SHFL + 62 SEL + 31
                    //      FADD of the shuffle version: the contraction warps are short of issue slots, not of LSU ----
                    float out = 0.f;
                    {
                        float *red = reinterpret_cast<float *>(smem + L.red_off) + cwarp * (4 * TI * WS_RED_PITCH) + lane;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q = 0; q < NP; ++q) {
                                float lo, hi;
                                unpack2(ap[q][b], lo, hi);
                                red[(4 * (2 * q) + b) * WS_RED_PITCH] = lo;
                                red[(4 * (2 * q + 1) + b) * WS_RED_PITCH] = hi;
                            }
                            if (ODD) red[(4 * (TI - 1) + b) * WS_RED_PITCH] = as[b];
                        }
                        __syncwarp();
                        if (lane < 4 * TI) {
                            const float4 *row = reinterpret_cast<const float4 *>(
                                reinterpret_cast<const float *>(smem + L.red_off) + cwarp * (4 * TI * WS_RED_PITCH) + lane * WS_RED_PITCH);
                            float4 v[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = row[k];
                            float s8[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) s8[k] = (v[k].x + v[k].y) + (v[k].z + v[k].w);
                            out = ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
                        }
                    }
#else
                    // ---- 32-lane reduce-scatter: row over lane bits 4..2, stimulus over bits 1..0 ----
                    float out;
                    {
                        const unsigned fullm = 0xffffffffu;
                        const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
                        float r8[8][4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q = 0; q < NP; ++q) unpack2(ap[q][b], r8[2 * q][b], r8[2 * q + 1][b]);
                            if (ODD) r8[TI - 1][b] = as[b];
#pragma unroll
                            for (int t = TI; t < 8; ++t) r8[t][b] = 0.f;
                        }
                        float r4[4][4], r2[2][4], r1[4], p2[2];
#pragma unroll
                        for (int t = 0; t < 4; ++t)
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const float send = u16 ? r8[t][b] : r8[4 + t][b];
                                const float keep = u16 ? r8[4 + t][b] : r8[t][b];
                                r4[t][b] = keep + __shfl_xor_sync(fullm, send, 16);
                            }
#pragma unroll
                        for (int t = 0; t < 2; ++t)
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const float send = u8 ? r4[t][b] : r4[2 + t][b];
                                const float keep = u8 ? r4[2 + t][b] : r4[t][b];
                                r2[t][b] = keep + __shfl_xor_sync(fullm, send, 8);
                            }
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const float send = u4 ? r2[0][b] : r2[1][b];
                            const float keep = u4 ? r2[1][b] : r2[0][b];
                            r1[b] = keep + __shfl_xor_sync(fullm, send, 4);
                        }
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            const float send = u2 ? r1[b] : r1[2 + b];
                            const float keep = u2 ? r1[2 + b] : r1[b];
                            p2[b] = keep + __shfl_xor_sync(fullm, send, 2);
                        }
                        {
                            const float send = u1 ? p2[0] : p2[1];
                            const float keep = u1 ? p2[1] : p2[0];
                            out = keep + __shfl_xor_sync(fullm, send, 1);
                        }
                    }
#endif
                    // lane = 4 * row + stimulus: hand the sum to the update warp
                    dvbuf[(h * CW + cwarp) * 32 + lane] = out;
                    __syncwarp();
                    if (lane == 0) mbar_arrive_release(smem_u32(&misc->dvfull[h][cwarp]));
#if SSN_WS_PROFILE
                    tc[3] += clock64() - c3;
#endif
                    bufbits ^= 1u << h;
                    ++it;
                };

                while (alive) {
                    if (alive & 1u) cstep(HalfC<0>{});
                    if (alive & 2u) cstep(HalfC<1>{});
                }
                // nobody may publish the next network's panels while a slower CTA still reads these
                cluster.sync();
            }
        }
#if SSN_WS_PROFILE
        if (a.dbg_out && ctid == 0)
            for (int q = 0; q < 4; ++q) a.dbg_out[rank * 8 + q] = tc[q];
#endif
    } else {
        // =====================================================================================
        // update warps: warp u serves contraction warps u, u + UW, .. (same scheduler); lane = 4 * row + stimulus slot owns one
        // output per served warp and stream
        // =====================================================================================
        reg_release<SH::REG_U>();
        const int u = warp - CW;                                    // update warp index
        const bool arming = u == 0 && lane == 0;
        const int my_t = lane >> 2, my_b = lane & 3;
        bool owner[G];
        int grow[G];
        unsigned xoff[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int lrow = ((u + g * UW)) * TI + my_t;                           // local row
            grow[g] = row_base + lrow;
            owner[g] = my_t < TI && lrow < rows_here;
            xoff[g] = 16u * (unsigned)grow[g] + 4u * (unsigned)my_b;            // slot of the global row
        }
        const unsigned foff = 16u * (unsigned)kpad + 4u * (unsigned)(rank * UW + u);
        unsigned ph = 0u, dvph = 0u;
#if SSN_WS_PROFILE
        long long tc[6] = {0, 0, 0, 0, 0, 0};
#endif
        // r_ref and v_ref of the owned outputs live in shared memory (they change only at refresh events and are
        // read once per sweep), r in registers
        auto ref_slot = [&](int h, int which, int g) -> double & {
            return refbuf[((h * 2 + which) * CW + (u + g * UW)) * 32 + lane];
        };
        // Publish: lanes 0..4*TI-1 send the new r - r_ref of their outputs, lane 28 the warp's flag word, to the
        // same panel offset in every CTA of the cluster: a plain store at home, st.async (remote store that
        // completes bytes on the destination's mbarrier) to the peers.  No proxy fence and no barrier among the
        // update warps.
        // (address deltas of the csize-1 peers in registers: a volatile shared-memory load per remote store, as in
        // the first version, serialised the whole loop behind its latency -- 31 % of the update warps' samples)
        unsigned dlt[MAX_CLUSTER - 1];
#pragma unroll
        for (int q = 0; q < MAX_CLUSTER - 1; ++q) dlt[q] = q < csize - 1 ? misc->pdelta[q + (q >= rank ? 1 : 0)] : 0u;
        // the flag lane rides in the instruction stream of output 0
        const bool send0 = owner[0] || lane == 28;
        const unsigned off0 = lane == 28 ? foff : xoff[0];
        auto publish = [&](int h, int nbuf, const unsigned (&bits)[G], unsigned flagword) {
            const unsigned base = (unsigned)((h * 2 + nbuf) * BUF_BYTES);
            const unsigned bar = full_bar(h, nbuf);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const bool send = g == 0 ? send0 : owner[g];
                if (send) {
                    const unsigned off = base + (g == 0 ? off0 : xoff[g]);
                    const unsigned val = (g == 0 && lane == 28) ? flagword : bits[g];
                    *reinterpret_cast<unsigned *>(smem + L.x_off + off) = val;
#pragma unroll
                    for (int q = 0; q < MAX_CLUSTER - 1; ++q)
                        if (q < csize - 1) st_async_u32(x_local + off + dlt[q], val, bar + dlt[q]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_release(bar);
        };
        for (;;) {
            cluster.sync();
            const int net = misc->next_net;
            if (net >= a.nz) break;
            const float *ext_net = a.ext + (size_t)net * a.ext_stride_z;

            {
                // stream h starts on half-panel h; a stream that finishes picks up the next one of this network
                int next_hp = min(2, n_hp);
                int hp0 = 0, hp1 = 1;                                        // half-panel of each stream
                unsigned done = empty_slots(0) | (empty_slots(1) << 4);
                unsigned force = a.r_init ? (~done & 0xffu) : 0u;
                unsigned alive = n_hp > 1 ? 3u : 1u;
                unsigned bufbits = 0u;
                int it0 = 1, it1 = 1;
                int my_status[2] = {1, 1}, my_iters[2] = {a.max_iter, a.max_iter};   // of slot lane & 3 of each stream

                // float64 state of the (stream, output) values a lane owns
                double sr[2][G];
                float sext[2][G];
                unsigned levels = 0u;                                        // ladder level, 4 bits per stream
                // (re)start stream h on half-panel hp: load the stimulus and the initial state of the owned outputs,
                // publish the initial panel r - r_ref (= r_init, refreshed at once, or 0) into buffer `buf`
                auto start_stream = [&](int h, int hp, int buf) {
                    const int sabs = 4 * hp + my_b;                          // stimulus of my slot
                    unsigned bits[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        double r0 = 0.0;
                        float e = 0.f;
                        if (owner[g] && sabs < a.nb) {
                            e = __ldg(ext_net + (size_t)sabs * dim + grow[g]);
                            if (a.r_init) r0 = (double)__ldg(a.r_init + ((size_t)net * a.nb + sabs) * dim + grow[g]);
                        }
                        sr[h][g] = r0; sext[h][g] = e;
                        ref_slot(h, 0, g) = 0.0;                             // r_ref
                        ref_slot(h, 1, g) = (double)e;                       // v_ref
                        bits[g] = __float_as_uint((float)r0);
                    }
                    levels &= ~(0xfu << (4 * h));
                    my_status[h] = 1; my_iters[h] = a.max_iter;
                    if (arming) mbar_arrive_expect_tx(full_bar(h, buf), tx_bytes);
                    publish(h, buf, bits, ws_mask(h) << 16);                 // flags: "big", no refresh yet
                };
                // the stream's half-panel is finished: write its results, then take the next half-panel (after a
                // cluster barrier, see the contraction role) or retire the stream
                auto u_finish = [&](int h) {
                    const int hp = h ? hp1 : hp0;
                    const int sabs = 4 * hp + my_b;
#pragma unroll
                    for (int g = 0; g < G; ++g)
                        if (owner[g] && sabs < a.nb) a.R[((size_t)net * a.nb + sabs) * dim + grow[g]] = (float)sr[h][g];
                    if (rank == 0 && u == 0 && lane < 4 && 4 * hp + lane < a.nb) {       // lane == slot for lanes 0..3
                        a.status[(size_t)net * a.nb + 4 * hp + lane] = my_status[h];
                        if (a.iters) a.iters[(size_t)net * a.nb + 4 * hp + lane] = my_iters[h];
                    }
                    if (next_hp < n_hp) {
                        cluster.sync();
                        const unsigned hm = ws_mask(h);
                        done = (done & ~hm) | (empty_slots(next_hp) << (4 * h));
                        force = (force & ~hm) | (a.r_init ? (~done & hm) : 0u);
                        (h ? hp1 : hp0) = next_hp;
                        (h ? it1 : it0) = 1;
                        bufbits ^= 1u << h;                 // the new initial panel goes to the other buffer
                        start_stream(h, next_hp, (bufbits >> h) & 1u);
                        ++next_hp;
                    } else {
                        alive &= ~(1u << h);
                    }
                };
                start_stream(0, 0, 0);
                if (n_hp > 1) start_stream(1, 1, 0);

                auto ustep = [&](auto hc) {
                    constexpr int h = decltype(hc)::value;
                    constexpr unsigned hm = ws_mask(h);
                    const int buf = (bufbits >> h) & 1u, nbuf = buf ^ 1;
                    int &it = h ? it1 : it0;
#if SSN_WS_PROFILE
                    long long c0 = clock64();
#endif
                    mbar_wait(full_bar(h, buf), (ph >> (2 * h + buf)) & 1u);
                    ph ^= 1u << (2 * h + buf);
#if SSN_WS_PROFILE
                    long long c1 = clock64(); tc[0] += c1 - c0;
#endif
                    const unsigned F = __reduce_or_sync(0xffffffffu, ws_flag_word<SH>(panel(h, buf), kpad, lane));
                    unsigned req, natural, conv_now, hard_now;
                    const bool go = advance(F, hm, it, done, force, req, natural, conv_now, hard_now);
                    if ((conv_now >> (4 * h + sid)) & 1u) { my_status[h] = 0; my_iters[h] = it - 1; }
                    if ((hard_now >> (4 * h + sid)) & 1u) { my_status[h] = 2; my_iters[h] = it - 1; }
                    if (!go) { u_finish(h); return; }
                    if (arming) mbar_arrive_expect_tx(full_bar(h, nbuf), tx_bytes);

                    const int st = ws_stim(h, my_b);
                    if (req) {
                        // ---- reference-point refresh: all-gather hi/lo of r, the contraction warps do the exact product ----
                        const bool slot_req = (req >> st) & 1u;
                        if (arming) mbar_arrive_expect_tx(smem_u32(&misc->xfull[h]), (unsigned)__popc(req) * (unsigned)dim * 8u);
                        if (slot_req) {
                            const unsigned xs = (unsigned)__popc(req & hm & ((1u << st) - 1u));
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                if (!owner[g]) continue;
                                const double ri = sr[h][g];
                                const float hi = (float)ri;
                                const float lo = (float)(ri - (double)hi);
                                const unsigned o = xe_local + 4u * (unsigned)(((h * 2) * 4 + xs) * kpad + grow[g]);
                                for (int p = 0; p < csize; ++p) {
                                    const unsigned pd = pdelta[p];
                                    const unsigned bar = smem_u32(&misc->xfull[h]) + pd;
                                    st_async_u32(o + pd, __float_as_uint(hi), bar);
                                    st_async_u32(o + 4u * (unsigned)(4 * kpad) + pd, __float_as_uint(lo), bar);
                                }
                            }
                        }
                        bar_sync_all(WS_BAR_REFRESH);
                        if (slot_req) {
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                if (!owner[g]) continue;
                                ref_slot(h, 1, g) = exbuf[(h * CW + (u + g * UW)) * 32 + lane] + (double)sext[h][g];
                                ref_slot(h, 0, g) = sr[h][g];
                            }
                            if ((natural >> st) & 1u) levels += 1u << (4 * h);      // next rung of the ladder
                        }
                    }
#if SSN_WS_PROFILE
                    long long c2 = clock64(); tc[1] += c2 - c1;
#endif
                    // ---- dv of the contraction warps this warp serves, float64 state update ----
                    const double tl = misc->tlevel[(levels >> (4 * h)) & 7u];
                    const bool slot_live = !((done >> st) & 1u);
                    unsigned word = 0u;
                    unsigned bits[G];
                    const unsigned dvpar = (dvph >> h) & 1u;
                    // all served contraction warps first: the G updates below are independent chains and interleave
#pragma unroll
                    for (int g = 0; g < G; ++g) mbar_wait(smem_u32(&misc->dvfull[h][(u + g * UW)]), dvpar);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float dv = dvbuf[(h * CW + (u + g * UW)) * 32 + lane];
                        const bool lv = owner[g] && slot_live;
                        const double vv = ref_slot(h, 1, g) + (double)dv;
                        bool rare;
#if SSN_WS_DIAG & 4
                        rare = false;
                        double fv = vv > 0.0 ? 0.5 * vv : 0.0;
#else
                        double fv = io_eval_common(a, tab, vv, rare);
                        if (rare) fv = io_eval_exact(a, vv);                       // beyond the tables: diverging networks
#endif
                        const double eps_own = grow[g] < N ? a.eps_E : a.eps_I;
                        const double d = lv ? (fv - sr[h][g]) * eps_own : 0.0;     // r_new - r_old
                        const double step = fabs(d);
                        const double r_cur = sr[h][g] + d;
                        if (lv && step >= a.atol) word |= 1u << st;
                        if (lv && r_cur >= a.r_hard) word |= 1u << (8 + st);
                        if (owner[g] && ((lv && step >= tl) || !(tl > 0.0))) word |= 1u << (16 + st);   // exhausted ladder never asks
                        sr[h][g] = r_cur;
                        bits[g] = __float_as_uint((float)(r_cur - ref_slot(h, 0, g)));
                    }
                    dvph ^= 1u << h;
#if SSN_WS_PROFILE
                    long long c4 = clock64(); tc[3] += c4 - c2;
#endif
                    word = __reduce_or_sync(0xffffffffu, word);
                    publish(h, nbuf, bits, word);
#if SSN_WS_PROFILE
                    tc[4] += clock64() - c4;
#endif
                    bufbits ^= 1u << h;
                    ++it;
                };

                while (alive) {
                    if (alive & 1u) ustep(HalfC<0>{});
                    if (alive & 2u) ustep(HalfC<1>{});
                }

                // nobody may publish the next network's panels while a slower CTA still reads these
                cluster.sync();
            }
        }
#if SSN_WS_PROFILE
        if (a.dbg_out && arming)
            for (int q = 0; q < 5; ++q) a.dbg_out[64 + rank * 8 + q] = tc[q];
#endif
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
typedef void (*WsKernel)(const RwArgs);
#ifndef SSN_WS_PF_A
#define SSN_WS_PF_A 2
#endif
#ifndef SSN_WS_PF_B
#define SSN_WS_PF_B 2
#endif
#ifndef SSN_WS_RC_B
#define SSN_WS_RC_B 144
#endif
#ifndef SSN_WS_RC_A
#define SSN_WS_RC_A 184
#endif
using ShapeA = WsShape<8, 8, 7, SSN_WS_RC_A, SSN_WS_PF_A>;             // 56 rows per CTA
#ifdef SSN_WS_SHAPE_B
using ShapeB = WsShape<12, 4, 6, SSN_WS_RC_B, SSN_WS_PF_B>;    // 72 rows per CTA (experiment, see the header)
#endif

struct WsPlan { WsKernel fn; int shape, nc, kpad, csize, rpc, smem, clusters, tab_nodes, cw, uw, ti; };

template <class SH>
static WsKernel pick_ws_kernel(int nc) {
    switch (nc) {
        case 2: return ssn_fp_ws_kernel<2, SH>;
        case 4: return ssn_fp_ws_kernel<4, SH>;
        case 7: return ssn_fp_ws_kernel<7, SH>;
        case 10: return ssn_fp_ws_kernel<10, SH>;
        case 13: return ssn_fp_ws_kernel<13, SH>;
        case 14: return ssn_fp_ws_kernel<14, SH>;
    }
    return nullptr;
}

static void ws_launch_config(const WsPlan &plan, int clusters, cudaStream_t stream, cudaLaunchConfig_t *cfg,
                             cudaLaunchAttribute *attr) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3(clusters * plan.csize, 1, 1);
    cfg->blockDim = dim3(WS_THREADS, 1, 1);
    cfg->dynamicSmemBytes = plan.smem;
    cfg->stream = stream;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
}

template <class SH>
static int plan_ws_shape(const ssn_solver &sv, int n_sites, int nz, int shape_id, WsPlan *plan) {
    const int dim = 2 * n_sites, rows = SH::ROWS;
    const int cands[] = {2, 4, 7, 10, 13, 14};
    plan->nc = 0;
    for (int nc : cands)
        if (32 * nc >= dim) { plan->nc = nc; break; }
    if (!plan->nc) return 1;                                   // too large: caller falls back
    plan->shape = shape_id; plan->cw = SH::CW; plan->uw = SH::UW; plan->ti = SH::TI;
    plan->kpad = 32 * plan->nc;
    plan->csize = (dim + rows - 1) / rows;
    if (plan->csize > MAX_CLUSTER) return 1;
    plan->rpc = (dim + plan->csize - 1) / plan->csize;
    if (plan->rpc * (plan->csize - 1) >= dim) return 1;
    plan->tab_nodes = rw_table_nodes(sv);
    plan->smem = ws_smem_layout<SH>(plan->kpad, n_sites, rw_table_bytes(plan->tab_nodes, rw_table2_nodes(sv))).total;
    int dev = 0, limit = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (plan->smem > limit) return 1;
    plan->fn = pick_ws_kernel<SH>(plan->nc);
    if (!plan->fn) return 1;
    SSN_CUDA(cudaFuncSetAttribute(plan->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    ws_launch_config(*plan, 1, nullptr, &cfg, attr);
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, plan->fn, &cfg) != cudaSuccess) { cudaGetLastError(); return 1; }
    if (max_clusters < 1) return 1;
    plan->clusters = nz > 0 ? std::min(max_clusters, nz) : max_clusters;
    return 0;
}

static int plan_ws(const ssn_solver &sv, int n_sites, int nz, WsPlan *plan) {
#ifdef SSN_WS_SHAPE_B
    const char *force = getenv("SSN_WS_SHAPE");
    if (force && (force[0] == 'B' || force[0] == 'b')) return plan_ws_shape<ShapeB>(sv, n_sites, nz, 1, plan);
#endif
    return plan_ws_shape<ShapeA>(sv, n_sites, nz, 0, plan);
}

int ws_occupancy(const ssn_solver &sv, int n_sites, int *cluster_size, int *resident_clusters) {
    WsPlan plan;
    int rc = plan_ws(sv, n_sites, 0, &plan);
    if (rc) return rc;
    if (cluster_size) *cluster_size = plan.csize;
    if (resident_clusters) *resident_clusters = plan.clusters;
    return 0;
}

int ws_kernel_name(const ssn_solver &sv, int n_sites, char *buf, int cap) {
    WsPlan plan;
    int rc = plan_ws(sv, n_sites, 0, &plan);
    if (rc) return rc;
    snprintf(buf, cap, "ssn_fp_ws_kernel<NC=%d,CW=%d,UW=%d,TI=%d>x%d", plan.nc, plan.cw, plan.uw, plan.ti, plan.csize);
    return 0;
}

// Returns 1 when the shape is outside this kernel's range (the caller then uses the shared-memory
// kernel), 0 on success, otherwise an error code.
int launch_fixed_point_ws(const ssn_solver &sv, int nz, int nb, int n_sites, int w_kind, const float *w,
                          const ssn_jds *jds, const float *ext, int ext_per_network, const float *r_init,
                          float *R, int *status, int *iters, int *counter, cudaStream_t stream) {
    WsPlan plan;
    int rc = plan_ws(sv, n_sites, nz, &plan);
    if (rc) return rc;
    RwArgs a = {};
    a.nz = nz; a.nb = nb; a.n_sites = n_sites; a.dim = 2 * n_sites;
    a.kpad = plan.kpad; a.csize = plan.csize; a.rpc = plan.rpc;
    a.w_kind = w_kind; a.w = w;
    if (w_kind == SSN_W_FROM_Z) {
        if (!jds) { set_error("SSN_W_FROM_Z needs jds"); return -1; }
        a.wc = make_weight_const(*jds, n_sites);
    }
    a.ext = ext; a.ext_stride_z = ext_per_network ? (long long)nb * 2 * n_sites : 0;
    a.r_init = r_init; a.R = R; a.status = status; a.iters = iters; a.work_counter = counter;
    rw_fill_solver_args(a, sv, plan.tab_nodes);

    SSN_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    ws_launch_config(plan, plan.clusters, stream, &cfg, attr);
#if SSN_WS_PROFILE
    SSN_CUDA(cudaMalloc(&a.dbg_out, 1280 * sizeof(long long)));
    SSN_CUDA(cudaMemset(a.dbg_out, 0, 1280 * sizeof(long long)));
#endif
    {
        KernelTimer kt("ssn_fp_ws_kernel", stream);
        SSN_CUDA(cudaLaunchKernelEx(&cfg, plan.fn, a));
    }
    count_launch();
#if SSN_WS_PROFILE
    {
        static long long h[1280];
        SSN_CUDA(cudaStreamSynchronize(stream));
        SSN_CUDA(cudaMemcpy(h, a.dbg_out, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[ssn ws] shape CW=%d UW=%d TI=%d NC=%d cluster %d resident clusters %d smem %d (cycles of one warp, all its networks)\n",
                plan.cw, plan.uw, plan.ti, plan.nc, plan.csize, plan.clusters, plan.smem);
        const char *cn[4] = {"wait", "-", "contract+flags(+refresh)", "reduce+handoff"};
        const char *un[5] = {"wait_panel", "flags/refresh", "-", "wait_dv+update", "publish"};
        for (int r = 0; r < plan.csize; r += (plan.csize > 1 ? plan.csize - 1 : 1)) {
            fprintf(stderr, "[ssn ws] rank %d contraction:", r);
            for (int q = 0; q < 4; ++q) fprintf(stderr, " %s=%lld", cn[q], h[r * 8 + q]);
            fprintf(stderr, "\n[ssn ws] rank %d update:", r);
            for (int q = 0; q < 5; ++q) fprintf(stderr, " %s=%lld", un[q], h[64 + r * 8 + q]);
            fprintf(stderr, "\n");
        }
        cudaFree(a.dbg_out);
    }
#endif
    return 0;
}

}  // namespace ssn
