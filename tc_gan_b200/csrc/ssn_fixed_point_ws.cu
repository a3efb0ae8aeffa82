// K1 v3: warp-specialised fixed-point solve with the weight matrix resident in registers.
//
// Same numerics, panel exchange and cluster shape as ssn_fixed_point_regw.cu (read its header
// first), but the sweep is no longer one serial chain executed by every warp in lockstep:
//
//   * 8 CONTRACTION warps hold the W tile (7 rows x NC columns per thread, packed pairs) and do
//     nothing but  wait panel -> FFMA2 contraction -> 32-lane reduce-scatter -> hand one dv per
//     lane to their update warp through shared memory;
//   * 8 UPDATE warps (warp u serves contraction warp u) own the float64 state (r, r_ref, v_ref in
//     registers), evaluate f from the tables, apply the Euler step and the stopping tests, and
//     publish the new r - r_ref (and the warp's flag word) to every CTA of the cluster with
//     st.async remote stores that complete bytes on the destination's mbarrier;
//   * the stimuli of a network run as TWO independent streams of half-panels (4 stimuli), each stream with
//     its own double-buffered panel and mbarriers.  While the update warps and the cluster exchange finish
//     sweep k of one stream, the contraction warps are already in sweep k of the other, so the float64
//     update, the publish and the DSMEM latency run under FMA work instead of after it.  A stream whose
//     half-panel has converged picks up the next half-panel of the network (nb > 8), so both streams stay
//     busy until the network runs out of stimuli; only then does the last stream run alone.
//
// Register budget: the CTA starts with 128 registers per thread (512 threads); the update warp
// groups release down to WS_REG_U and the contraction warp groups grow to WS_REG_C with
// setmaxnreg (256 * 184 + 256 * 72 = the CTA's 65536), which is what lets a 98-register W tile
// coexist with a second set of warps.  (This file must NOT be compiled with -rdc: ptxas ignores
// setmaxnreg in relocatable device code.)
//
// Stopping rule and error codes follow tc_gan/ext/ssnode.c:84-102 exactly as in the regw kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "ssn_ws_common.cuh"
#include "ssn_launch.h"

#ifndef SSN_WS_PROFILE
#define SSN_WS_PROFILE 0
#endif
// SSN_WS_PROFILE=2 additionally records a clock64 timeline of sweeps 100..107 of network 0 (CTA rank 0)
#if SSN_WS_PROFILE == 2
#define WS_TRACE(on, ev) do { if ((on) && net == 0 && it >= 100 && it < 108) a.dbg_out[128 + rank * 128 + ((it - 100) * 2 + h) * 8 + (ev)] = clock64(); } while (0)
#else
#define WS_TRACE(on, ev) do { } while (0)
#endif
// tuning switches (kept for the measurements quoted in DESIGN.md)
#ifndef SSN_WS_SPEC
#define SSN_WS_SPEC 1        // contract speculatively, evaluate the flag word afterwards
#endif
#ifndef SSN_WS_PF
#define SSN_WS_PF 2          // panel columns loaded this many columns ahead of use
#endif
#ifndef SSN_WS_REG_C
#define SSN_WS_REG_C 184
#define SSN_WS_REG_U 72
#endif
#ifndef SSN_WS_U_FIRST
#define SSN_WS_U_FIRST 0     // 1: update warps take the lowest warp ids
#endif

namespace ssn {

constexpr int WS_CW = 8;                                // contraction warps
constexpr int WS_UW = 8;                                // update warps; warp u serves contraction warp u
constexpr int WS_THREADS = 32 * (WS_CW + WS_UW);
constexpr int WS_TI = 7;                                // rows per contraction warp
constexpr int WS_NP = WS_TI / 2;
// State panel of a stream (one buffer): per source CTA a slab of `slab_slots` 16-byte slots: the float4 (four
// stimuli) of its WS_ROWS local rows, then the flag words of its update warps (two slots).
// slab_slots = rows per CTA (mod 8) and >= WS_SLAB_USED keeps the LDS.128 of 8 consecutive columns on 8 distinct
// 16-byte bank groups also where they straddle two CTAs.  The slot after the last slab is always zero.
constexpr int WS_ROWS = WS_CW * WS_TI;                  // 56
constexpr int WS_SLAB_USED = WS_ROWS + WS_UW / 4;       // rows + flag slots
constexpr int WS_SLAB_MAX = WS_SLAB_USED + 7;
constexpr int WS_BUF_BYTES = (MAX_CLUSTER * WS_SLAB_MAX + 1) * 16;
__host__ __device__ constexpr int ws_slab_slots(int rpc) { return WS_SLAB_USED + (((rpc - WS_SLAB_USED) % 8) + 8) % 8; }
constexpr int WS_REG_C = SSN_WS_REG_C, WS_REG_U = SSN_WS_REG_U;            // 256 C + 256 U <= 512 * 128 (the CTA's allocation)
constexpr int WS_BAR_REFRESH = 1;                       // named barrier used by refresh events (all threads)

struct WsMisc {
    unsigned long long full[2][2];          // [half][buffer]: panel of the next sweep complete
    unsigned long long xfull[2];            // [half]: hi/lo columns of a refresh event complete
    unsigned long long dvfull[2][WS_CW];    // [half][contraction warp]: dv handed over
    double tlevel[8];                       // refresh ladder thresholds by level, 0 = exhausted
    unsigned pdelta[MAX_CLUSTER];
    int next_net;
};
static_assert(sizeof(WsMisc) <= 512, "misc block");

struct WsSmem {
    int x_off, xe_off, tab_off, gtab_off, dv_off, ex_off, misc_off, total;
};
__host__ __device__ inline WsSmem ws_smem_layout(int kpad, int n_sites, int tab_bytes) {
    WsSmem L;
    int o = 0;
    L.x_off = o;    o += 4 * WS_BUF_BYTES;                  // [half][buffer]
    L.xe_off = o;   o += 2 * 2 * 4 * kpad * 4;              // [half][hi, lo][4 stimuli][kpad]
    L.tab_off = o;  o += tab_bytes;                       // Taylor tables of f
    L.gtab_off = o; o += ((4 * n_sites * 4 + 15) / 16) * 16;
    L.dv_off = o;   o += 2 * WS_CW * 32 * 4;                // [half][contraction warp][lane] float
    L.ex_off = o;   o += 2 * WS_CW * 32 * 8;                // [half][contraction warp][lane] double
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

// OR of the flag words of all blocks of a panel buffer (absent blocks stay zero)
__device__ __forceinline__ unsigned ws_flags(const unsigned char *buf_base, int slab_slots, int lane) {
    const uint2 f = *reinterpret_cast<const uint2 *>(buf_base + ((lane >> 2) * slab_slots + WS_ROWS) * 16 + (lane & 3) * 8);
    return __reduce_or_sync(0xffffffffu, f.x | f.y);
}

template <int NC>
__global__ void __launch_bounds__(WS_THREADS, 1) ssn_fp_ws_kernel(const RwArgs a) {
    static_assert(MAX_CLUSTER * WS_UW == 64, "two flag words per lane");
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int csize = a.csize, dim = a.dim, kpad = a.kpad, rpc = a.rpc, N = a.n_sites;
    const int slab_slots = ws_slab_slots(rpc);
    const WsSmem L = ws_smem_layout(kpad, N, rw_table_bytes(a.tab_nodes, a.tab2_nodes));
    float *xe = reinterpret_cast<float *>(smem + L.xe_off);
    double *tab = reinterpret_cast<double *>(smem + L.tab_off);
    float *gtab = reinterpret_cast<float *>(smem + L.gtab_off);
    float *dvbuf = reinterpret_cast<float *>(smem + L.dv_off);
    double *exbuf = reinterpret_cast<double *>(smem + L.ex_off);
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
            mbar_init(smem_u32(&misc->full[h][0]), 1 + WS_UW);     // the arming thread + one release-arrive per update warp
            mbar_init(smem_u32(&misc->full[h][1]), 1 + WS_UW);
            mbar_init(smem_u32(&misc->xfull[h]), 1);
            for (int w = 0; w < WS_CW; ++w) mbar_init(smem_u32(&misc->dvfull[h][w]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 4 * WS_BUF_BYTES / 4; i += WS_THREADS) reinterpret_cast<float *>(smem + L.x_off)[i] = 0.f;
    for (int i = tid; i < 2 * 2 * 4 * kpad; i += WS_THREADS) xe[i] = 0.f;
    if (a.w_kind == SSN_W_FROM_Z) build_profile_table(a.wc, N, gtab, tid, WS_THREADS);
    build_io_tables(a, tab, tid, WS_THREADS);
    cluster.sync();

    const volatile unsigned *pdelta = misc->pdelta;
    // bytes arriving from the peers per panel: 16 per row they own + their update warps' flag words
    const unsigned tx_bytes = (unsigned)((dim - rows_here) * 16 + (csize - 1) * WS_UW * 4);
    const int n_hp = (a.nb + 3) / 4;                        // half-panels (4 stimuli) of a network: the units the streams pull
    auto full_bar = [&](int h, int b) { return smem_u32(&misc->full[h][b]); };
    auto panel = [&](int h, int b) { return smem + L.x_off + (h * 2 + b) * WS_BUF_BYTES; };

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

    constexpr bool U_FIRST = SSN_WS_U_FIRST != 0;
    const bool contraction_role = U_FIRST ? warp >= WS_UW : warp < WS_CW;
    const int cwarp = U_FIRST ? warp - WS_UW : warp;                 // index among the contraction warps
    const int ctid = tid - (U_FIRST ? 32 * WS_UW : 0);               // thread index among them
    if (contraction_role) {
        // =====================================================================================
        // contraction warps
        // =====================================================================================
        reg_grow<WS_REG_C>();
        const int row0 = cwarp * WS_TI;
        // slot (16-byte unit inside a buffer) of panel column j = c*32 + lane, two per register
        unsigned colslot[(NC + 1) / 2];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = c * 32 + lane;
            unsigned slot = (unsigned)(MAX_CLUSTER * slab_slots);          // the zero slot
            if (j < dim) {
                const int cta = j / rpc, lr = j - cta * rpc;
                slot = (unsigned)(cta * slab_slots + lr);
            }
            if (c & 1) colslot[c / 2] |= slot << 16; else colslot[c / 2] = slot;
        }
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

            // ---- W tile -> registers (all loads issued before any is consumed) ----
            unsigned long long wp[WS_NP][NC];
            float ws[NC];
            {
                const float *src = a.w + (size_t)net * dim * dim + (size_t)(row_base + row0) * dim + lane;
                float zv[WS_TI][NC];
#pragma unroll
                for (int t = 0; t < WS_TI; ++t)
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        zv[t][c] = (row0 + t < rows_here && c * 32 + lane < dim) ? __ldg(src + (size_t)t * dim + c * 32) : 0.f;
                if (a.w_kind == SSN_W_FROM_Z) {
#pragma unroll
                    for (int t = 0; t < WS_TI; ++t) {
                        const int i = row_base + row0 + t;
                        const int ah = i >= N, ii = i - ah * N;
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            const int j = c * 32 + lane;
                            const int bh = j >= N, ab = ah * 2 + bh;
                            int d = ii - (j - bh * N);
                            d = d < 0 ? -d : d;
                            const bool ok = row0 + t < rows_here && j < dim;
                            zv[t][c] = ok ? gtab[ab * N + min(d, N - 1)] * fmaf(a.wc.sD[ab], zv[t][c], a.wc.sJ[ab]) : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) {
#pragma unroll
                    for (int q = 0; q < WS_NP; ++q) wp[q][c] = pack2(zv[2 * q][c], zv[2 * q + 1][c]);
                    ws[c] = zv[WS_TI - 1][c];
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
                    WS_TRACE(cwarp == 0 && lane == 0, 0);
                    unsigned char *xb = panel(h, buf);
                    // the flag word is loaded now and looked at after the contraction: finishing and refresh
                    // events are rare, so the contraction runs speculatively under the latency of the flag logic
                    const uint2 fword2 = *reinterpret_cast<const uint2 *>(xb + ((lane >> 2) * slab_slots + WS_ROWS) * 16 + (lane & 3) * 8);
                    const unsigned fword = fword2.x | fword2.y;
                    unsigned long long ap[WS_NP][4];
                    float as[4];
                    // ---- contraction of the half: dv = W * fl32(r - r_ref), four stimuli per LDS.128,
                    //      panel columns fetched SSN_WS_PF columns ahead of their use ----
                    auto contract = [&]() {
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q = 0; q < WS_NP; ++q) ap[q][b] = 0ull;
                            as[b] = 0.f;
                        }
                        const float4 *Xq = reinterpret_cast<const float4 *>(xb);
                        constexpr int PF = SSN_WS_PF;
                        float4 xq[PF + 1];
#pragma unroll
                        for (int c = 0; c < PF && c < NC; ++c)
                            xq[c] = Xq[(c & 1) ? (colslot[c / 2] >> 16) : (colslot[c / 2] & 0xffffu)];
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            if (c + PF < NC) {
                                const int cn = c + PF;
                                xq[cn % (PF + 1)] = Xq[(cn & 1) ? (colslot[cn / 2] >> 16) : (colslot[cn / 2] & 0xffffu)];
                            }
                            const float4 x4 = xq[c % (PF + 1)];
                            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
#pragma unroll
                                for (int q = 0; q < WS_NP; ++q) ffma2(ap[q][b], wp[q][c], xv[b]);
                                as[b] = fmaf(ws[c], xv[b], as[b]);
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
                                double accd[WS_TI];
                                float accf[WS_TI];
#pragma unroll
                                for (int t = 0; t < WS_TI; ++t) { accd[t] = 0.0; accf[t] = 0.f; }
#pragma unroll
                                for (int c = 0; c < NC; ++c) {
                                    const double hv = (double)xh[c * 32 + lane];
                                    const float lv = xl[c * 32 + lane];
                                    float wv[WS_TI];
#pragma unroll
                                    for (int q = 0; q < WS_NP; ++q) unpack2(wp[q][c], wv[2 * q], wv[2 * q + 1]);
                                    wv[WS_TI - 1] = ws[c];
#pragma unroll
                                    for (int t = 0; t < WS_TI; ++t) {
                                        accd[t] = fma((double)wv[t], hv, accd[t]);          // exact products, fp64 sum
                                        accf[t] = fmaf(wv[t], lv, accf[t]);
                                    }
                                }
                                double mine = 0.0;
#pragma unroll
                                for (int t = 0; t < WS_TI; ++t) {
                                    double v = accd[t] + (double)accf[t];
#pragma unroll
                                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                    mine = (t == (lane >> 2)) ? v : mine;
                                }
                                if ((lane & 3) == b) exbuf[(h * WS_CW + cwarp) * 32 + lane] = mine;
                                // r - r_ref is now zero for this stimulus on every row of every CTA
                                float *col = reinterpret_cast<float *>(xb) + b;
                                for (int q = ctid; q < MAX_CLUSTER * WS_ROWS; q += 32 * WS_CW) {
                                    const int cta = q / WS_ROWS, lr = q - cta * WS_ROWS;
                                    col[4 * (cta * slab_slots + lr)] = 0.f;
                                }
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
                    WS_TRACE(cwarp == 0 && lane == 0, 1);
                    // ---- 32-lane reduce-scatter: row over lane bits 4..2, stimulus over bits 1..0 ----
                    float out;
                    {
                        const unsigned fullm = 0xffffffffu;
                        const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
                        float r8[8][4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q = 0; q < WS_NP; ++q) unpack2(ap[q][b], r8[2 * q][b], r8[2 * q + 1][b]);
                            r8[WS_TI - 1][b] = as[b];
#pragma unroll
                            for (int t = WS_TI; t < 8; ++t) r8[t][b] = 0.f;
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
                    // lane = 4 * row + stimulus: hand the sum to the update warp
                    dvbuf[(h * WS_CW + cwarp) * 32 + lane] = out;
                    __syncwarp();
                    if (lane == 0) mbar_arrive_release(smem_u32(&misc->dvfull[h][cwarp]));
                    WS_TRACE(cwarp == 0 && lane == 0, 2);
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
        // update warps: warp u serves contraction warp u; lane = 4 * row + stimulus slot owns one output per stream
        // =====================================================================================
        reg_release<WS_REG_U>();
        const int u = U_FIRST ? warp : warp - WS_CW;                // update warp index
        const bool arming = u == 0 && lane == 0;
        const int my_t = lane >> 2, my_b = lane & 3;
        const int lrow = u * WS_TI + my_t;                          // local row
        const int grow = row_base + lrow;
        const bool owner = my_t < WS_TI && lrow < rows_here;
        const double eps_own = grow < N ? a.eps_E : a.eps_I;
        const unsigned slab = (unsigned)(rank * slab_slots * 16);               // this CTA's slab in a panel buffer
        const unsigned xoff = slab + 16u * (unsigned)lrow + 4u * (unsigned)my_b;
        const unsigned foff = slab + 16u * WS_ROWS + 4u * (unsigned)u;
        unsigned ph = 0u, dvph = 0u;
#if SSN_WS_PROFILE
        long long tc[6] = {0, 0, 0, 0, 0, 0};
#endif
        // Publish one value per lane (lanes 0..27: the new r - r_ref of their output; lane 28: the warp's flag word)
        // to the same panel offset in every CTA of the cluster: a plain store at home, st.async (remote store that
        // completes bytes on the destination's mbarrier) to the peers.  No proxy fence and no barrier among the
        // update warps, unlike a cp.async.bulk of the slab (measured equal in throughput, ~500 cycles slower per
        // exchange when a single stream is left).
        const bool sender = owner || lane == 28;
        const unsigned my_off = lane == 28 ? foff : xoff;
        auto publish = [&](int h, int nbuf, unsigned bits) {
            const unsigned off = (unsigned)((h * 2 + nbuf) * WS_BUF_BYTES) + my_off;
            if (sender) {
                *reinterpret_cast<unsigned *>(smem + L.x_off + off) = bits;
#pragma unroll
                for (int q = 0; q < MAX_CLUSTER - 1; ++q) {
                    if (q < csize - 1) {
                        const unsigned dlt = misc->pdelta[q + (q >= rank ? 1 : 0)];       // the csize-1 peers
                        st_async_u32(x_local + off + dlt, bits, full_bar(h, nbuf) + dlt);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_release(full_bar(h, nbuf));
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
                double sr[2], srref[2], svref[2];
                float sext[2];
                unsigned levels = 0u;                                        // ladder level, 4 bits per stream
                // (re)start stream h on half-panel hp: load the stimulus and the initial state of the owned output,
                // publish the initial panel r - r_ref (= r_init, refreshed at once, or 0) into buffer `buf`
                auto start_stream = [&](int h, int hp, int buf) {
                    double r0 = 0.0;
                    float e = 0.f;
                    const int sabs = 4 * hp + my_b;                          // stimulus of my slot
                    if (owner && sabs < a.nb) {
                        e = __ldg(ext_net + (size_t)sabs * dim + grow);
                        if (a.r_init) r0 = (double)__ldg(a.r_init + ((size_t)net * a.nb + sabs) * dim + grow);
                    }
                    sr[h] = r0; srref[h] = 0.0; svref[h] = (double)e; sext[h] = e;
                    levels &= ~(0xfu << (4 * h));
                    my_status[h] = 1; my_iters[h] = a.max_iter;
                    if (arming) mbar_arrive_expect_tx(full_bar(h, buf), tx_bytes);
                    publish(h, buf, lane == 28 ? ws_mask(h) << 16 : __float_as_uint((float)r0));     // flags: "big", no refresh yet
                };
                // the stream's half-panel is finished: write its results, then take the next half-panel (after a
                // cluster barrier, see the contraction role) or retire the stream
                auto u_finish = [&](int h) {
                    const int hp = h ? hp1 : hp0;
                    const int sabs = 4 * hp + my_b;
                    if (owner && sabs < a.nb) a.R[((size_t)net * a.nb + sabs) * dim + grow] = (float)sr[h];
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
                    WS_TRACE(u == 0 && lane == 0, 3);
                    const unsigned F = ws_flags(panel(h, buf), slab_slots, lane);
                    unsigned req, natural, conv_now, hard_now;
                    const bool go = advance(F, hm, it, done, force, req, natural, conv_now, hard_now);
                    if ((conv_now >> (4 * h + sid)) & 1u) { my_status[h] = 0; my_iters[h] = it - 1; }
                    if ((hard_now >> (4 * h + sid)) & 1u) { my_status[h] = 2; my_iters[h] = it - 1; }
                    if (!go) { u_finish(h); return; }
                    if (arming) mbar_arrive_expect_tx(full_bar(h, nbuf), tx_bytes);

                    const int st = ws_stim(h, my_b);
                    if (req) {
                        // ---- reference-point refresh: all-gather hi/lo of r, the contraction warps do the exact product ----
                        const bool mine_req = owner && ((req >> st) & 1u);
                        if (arming) mbar_arrive_expect_tx(smem_u32(&misc->xfull[h]), (unsigned)__popc(req) * (unsigned)dim * 8u);
                        if (mine_req) {
                            const unsigned xs = (unsigned)__popc(req & hm & ((1u << st) - 1u));
                            const double ri = sr[h];
                            const float hi = (float)ri;
                            const float lo = (float)(ri - (double)hi);
                            const unsigned o = xe_local + 4u * (unsigned)(((h * 2) * 4 + xs) * kpad + grow);
                            for (int p = 0; p < csize; ++p) {
                                const unsigned bar = smem_u32(&misc->xfull[h]) + pdelta[p];
                                st_async_u32(o + pdelta[p], __float_as_uint(hi), bar);
                                st_async_u32(o + 4u * (unsigned)(4 * kpad) + pdelta[p], __float_as_uint(lo), bar);
                            }
                        }
                        bar_sync_all(WS_BAR_REFRESH);
                        if (mine_req) {
                            svref[h] = exbuf[(h * WS_CW + u) * 32 + lane] + (double)sext[h];
                            srref[h] = sr[h];
                            if ((natural >> st) & 1u) levels += 1u << (4 * h);      // next rung of the ladder
                        }
                    }
#if SSN_WS_PROFILE
                    long long c2 = clock64(); tc[1] += c2 - c1;
#endif
                    // ---- dv of the contraction warp this warp serves ----
                    mbar_wait(smem_u32(&misc->dvfull[h][u]), (dvph >> h) & 1u);
                    dvph ^= 1u << h;
                    WS_TRACE(u == 0 && lane == 0, 4);
#if SSN_WS_PROFILE
                    long long c3 = clock64(); tc[2] += c3 - c2;
#endif
                    const float dv = dvbuf[(h * WS_CW + u) * 32 + lane];

                    // ---- float64 state update ----
                    unsigned word = 0u;
                    const bool lv = owner && !((done >> st) & 1u);
                    const double vv = svref[h] + (double)dv;
                    bool rare;
                    double fv = io_eval_common(a, tab, vv, rare);
                    if (rare) fv = io_eval_exact(a, vv);                       // beyond the tables: diverging networks
                    const double tl = misc->tlevel[(levels >> (4 * h)) & 7u];
                    const double d = lv ? (fv - sr[h]) * eps_own : 0.0;         // r_new - r_old
                    const double step = fabs(d);
                    const double r_cur = sr[h] + d;
                    if (lv && step >= a.atol) word |= 1u << st;
                    if (lv && r_cur >= a.r_hard) word |= 1u << (8 + st);
                    if (owner && ((lv && step >= tl) || !(tl > 0.0))) word |= 1u << (16 + st);   // exhausted ladder never asks
                    sr[h] = r_cur;
                    const float xn = (float)(r_cur - srref[h]);
#if SSN_WS_PROFILE
                    long long c4 = clock64(); tc[3] += c4 - c3;
#endif
                    WS_TRACE(u == 0 && lane == 0, 5);
                    word = __reduce_or_sync(0xffffffffu, word);
                    publish(h, nbuf, lane == 28 ? word : __float_as_uint(xn));
#if SSN_WS_PROFILE
                    tc[4] += clock64() - c4;
#endif
                    WS_TRACE(u == 0 && lane == 0, 6);
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
static WsKernel pick_ws_kernel(int nc) {
    switch (nc) {
        case 2: return ssn_fp_ws_kernel<2>;
        case 4: return ssn_fp_ws_kernel<4>;
        case 7: return ssn_fp_ws_kernel<7>;
        case 10: return ssn_fp_ws_kernel<10>;
        case 14: return ssn_fp_ws_kernel<14>;
    }
    return nullptr;
}

struct WsPlan { WsKernel fn; int nc, kpad, csize, rpc, smem, clusters, tab_nodes; };

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

static int plan_ws(const ssn_solver &sv, int n_sites, int nz, WsPlan *plan) {
    const int dim = 2 * n_sites, rows = WS_TI * WS_CW;
    const int cands[] = {2, 4, 7, 10, 14};
    plan->nc = 0;
    for (int nc : cands)
        if (32 * nc >= dim) { plan->nc = nc; break; }
    if (!plan->nc) return 1;                                   // too large: caller falls back
    plan->kpad = 32 * plan->nc;
    plan->csize = (dim + rows - 1) / rows;
    if (plan->csize > MAX_CLUSTER) return 1;
    plan->rpc = (dim + plan->csize - 1) / plan->csize;
    if (plan->rpc * (plan->csize - 1) >= dim) return 1;
    plan->tab_nodes = rw_table_nodes(sv);
    plan->smem = ws_smem_layout(plan->kpad, n_sites, rw_table_bytes(plan->tab_nodes, rw_table2_nodes(sv))).total;
    int dev = 0, limit = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (plan->smem > limit) return 1;
    plan->fn = pick_ws_kernel(plan->nc);
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
    snprintf(buf, cap, "ssn_fp_ws_kernel<NC=%d,CW=%d,UW=%d,TI=%d>x%d", plan.nc, WS_CW, WS_UW, WS_TI, plan.csize);
    return 0;
}

// Returns 1 when the shape is outside this kernel's range (the caller then uses the lockstep
// register kernel), 0 on success, otherwise an error code.
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
        fprintf(stderr, "[ssn ws] cluster %d resident clusters %d smem %d (cycles of one warp, all its networks)\n",
                plan.csize, plan.clusters, plan.smem);
        const char *cn[4] = {"wait", "-", "contract+flags(+refresh)", "reduce+handoff"};
        const char *un[5] = {"wait_panel", "flags/refresh", "wait_dv", "update", "publish"};
        for (int r = 0; r < plan.csize; r += (plan.csize > 1 ? plan.csize - 1 : 1)) {
            fprintf(stderr, "[ssn ws] rank %d contraction:", r);
            for (int q = 0; q < 4; ++q) fprintf(stderr, " %s=%lld", cn[q], h[r * 8 + q]);
            fprintf(stderr, "\n[ssn ws] rank %d update:", r);
            for (int q = 0; q < 5; ++q) fprintf(stderr, " %s=%lld", un[q], h[64 + r * 8 + q]);
            fprintf(stderr, "\n");
        }
#if SSN_WS_PROFILE == 2
        for (int r = 0; r < plan.csize; ++r) {
            // events: 0 C:panel, 1 C:contracted, 2 C:handed, 3 U:panel, 4 U:dv, 5 U:updated, 6 U:published  (per sweep 100..107, stream)
            const long long *t = h + 128 + r * 128;
            double c_flags_contract = 0, c_reduce = 0, handoff = 0, u_update = 0, u_publish = 0, exch = 0, c_idle = 0, period = 0;
            int n = 0;
            for (int s2 = 2; s2 < 14; ++s2) {                 // (it, h) pairs with a predecessor and a successor
                const long long *e = t + s2 * 8, *prev = t + (s2 - 1) * 8, *next = t + (s2 + 2) * 8;
                if (!e[0] || !next[0] || !prev[2]) continue;
                c_flags_contract += e[1] - e[0]; c_reduce += e[2] - e[1]; handoff += e[4] - e[2];
                u_update += e[5] - e[4]; u_publish += e[6] - e[5]; exch += next[0] - e[6];
                c_idle += e[0] - prev[2]; period += next[0] - e[0];
                ++n;
            }
            if (n)
                fprintf(stderr, "[ssn ws] rank %d: C flags+contract %.0f, C reduce+handoff %.0f, dv latency %.0f, U update %.0f, U publish %.0f, "
                        "publish->next panel %.0f, C idle before the step %.0f, sweep period %.0f\n", r, c_flags_contract / n,
                        c_reduce / n, handoff / n, u_update / n, u_publish / n, exch / n, c_idle / n, period / n);
        }
#endif
        cudaFree(a.dbg_out);
    }
#endif
    return 0;
}

}  // namespace ssn
