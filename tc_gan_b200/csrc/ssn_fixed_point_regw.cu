// K1 v2: fixed-point solve with the weight matrix resident in REGISTERS.
//
// A cluster of csize = ceil(2N / 56) CTAs (8 for 2N = 402) owns one network.  CTA `rank`
// owns rows [rank*rpc, rank*rpc + rpc); warp w of its 8 warps owns 7 of those rows and
// lane l the columns {l, l+32, l+64, ...}, so each thread keeps a 7 x NC tile of W in
// registers for the whole life of the network (NC = kpad/32 <= 14: 98 registers at
// 2N = 402) and the sweep loop issues NO shared-memory loads of W: per column two
// conflict-free LDS.128 of the 8-stimulus state panel feed 56 FFMA.
//
// Exchange: every (row, stimulus) owner publishes its new state to all CTAs of the
// cluster with st.async (remote store that completes bytes on the destination's
// mbarrier); a CTA starts its next sweep when its own mbarrier phase completes.  There
// is no barrier.cluster and no __syncthreads in the sweep loop.
//
// Numerics ("reference-point iteration"): per stimulus the kernel iterates on
// dr = r - r_ref with  v = v_ref + W * fl32(r - r_ref)  (FP32 FFMA on a small vector),
// f evaluated in float64 from a cubic table, and the state update in float64.  r_ref
// starts at the initial state (v_ref = I exactly for r_init = 0) and is refreshed -- an
// exact W * r_ref with float64 accumulation of exact fp32 x fp32 products for one
// panel column -- every time max|dr| has shrunk 64-fold, so the contraction error stays
// ~1e-7 RELATIVE to the remaining distance to the fixed point.  The sweep at which
// |r_new - r_old| < atol first holds then matches the float64 reference solver
// (tc_gan/ext/ssnode.c:84-96) instead of jittering by tens of sweeps.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "ssn_regw_common.cuh"
#include "ssn_launch.h"

namespace ssn {

constexpr int RW_MAXC = 16;                            // largest cluster (non-portable size, 4-warp CTAs)
constexpr int RW_BLOCKS = 80;                          // csize * warps <= 80 panel blocks
constexpr int RW_XE = 4;                               // stimuli refreshed per event
// State panel: one block per (source CTA, warp): 7 float4 rows of stimuli 0..3, 7 float4 rows of
// stimuli 4..7, one 16-byte slot whose first word carries the warp's flags.  A warp publishes its
// block to a peer CTA with ONE cp.async.bulk (one mbarrier transaction per block, not per element).
// (TI = rows per warp: 7 -> 15 slots, 240 B per block; the slot after the last block is always zero)
__host__ __device__ constexpr int rw_blk_slots(int ti) { return 2 * ti + 1; }
__host__ __device__ constexpr int rw_buf_bytes(int ti) { return (RW_BLOCKS * rw_blk_slots(ti) + 1) * 16; }
struct RwSmem {
    int x_off, xe_off, tab_off, gtab_off, state_off, misc_off, total;
};
// per-thread float64 state of the owner lanes, kept in shared memory so that the sweep loop's
// registers hold only the W tile and the accumulators: r, r_ref, v_ref as [2][threads] doubles, ext as [2][threads] floats
__host__ __device__ constexpr int rw_state_bytes(int nt) { return 3 * 2 * nt * 8 + 2 * nt * 4; }
struct RwMisc {
    unsigned long long full[2], xfull;
    double tlevel[8];                        // refresh ladder: thresholds on max|dr| by level, 0 = exhausted
    unsigned pdelta[RW_MAXC];
    int next_net;
};
__host__ __device__ inline RwSmem rw_smem_layout(int kpad, int n_sites, int tab_bytes, int nt, int ti) {
    RwSmem L;
    int o = 0;
    L.x_off = o;    o += 2 * rw_buf_bytes(ti);         // [buf][source CTA][warp] blocks + a zero slot
    L.xe_off = o;   o += 2 * RW_XE * kpad * 4;         // exact-pass columns: hi[4][kpad], lo[4][kpad]
    L.tab_off = o;  o += tab_bytes;                       // Taylor tables of f
    L.gtab_off = o; o += ((4 * n_sites * 4 + 15) / 16) * 16;
    L.state_off = o; o += rw_state_bytes(nt);
    L.misc_off = o; o += 1024;
    L.total = o;
    return L;
}

template <int NC, int NW, int TI>
__global__ void __launch_bounds__(32 * NW, NW == 4 ? 2 : 1) ssn_fp_regw_kernel(const RwArgs a) {
    constexpr int RW_WARPS = NW, RW_THREADS = 32 * NW, RW_TI = TI, NP = TI / 2;
    constexpr bool ODD = (TI & 1) != 0;                   // a single (unpaired) last row
    constexpr int RW_BLK_SLOTS = rw_blk_slots(TI), RW_BLK_BYTES = RW_BLK_SLOTS * 16;
    constexpr int RW_ZERO_SLOT = RW_BLOCKS * RW_BLK_SLOTS, RW_BUF_BYTES = rw_buf_bytes(TI);
    static_assert(TI >= 2 && TI <= 8, "rows per warp");
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int csize = a.csize, dim = a.dim, kpad = a.kpad, rpc = a.rpc, N = a.n_sites;
    const RwSmem L = rw_smem_layout(kpad, N, rw_table_bytes(a.tab_nodes, a.tab2_nodes), RW_THREADS, TI);
    float *Xf = reinterpret_cast<float *>(smem + L.x_off);
    float *xe = reinterpret_cast<float *>(smem + L.xe_off);
    double *tab = reinterpret_cast<double *>(smem + L.tab_off);
    float *gtab = reinterpret_cast<float *>(smem + L.gtab_off);
    RwMisc *misc = reinterpret_cast<RwMisc *>(smem + L.misc_off);
    double *sR = reinterpret_cast<double *>(smem + L.state_off) + threadIdx.x;        // [i][256]
    double *sRref = sR + 2 * RW_THREADS, *sVref = sR + 4 * RW_THREADS;
    float *sExt = reinterpret_cast<float *>(smem + L.state_off + 3 * 2 * RW_THREADS * 8) + threadIdx.x;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = rank * rpc;
    const int rows_here = max(0, min(rpc, dim - row_base));
    const int row0 = warp * RW_TI;                                  // first local row of this warp

    // ownership after the 32-lane reduce-scatter: lane bit 4 -> stimulus half h (stimuli 4h..4h+3),
    // bits 3..1 -> row t of the warp's seven, bit 0 -> which two of the four stimuli of the half.
    const int my_half = lane >> 4;
    const int my_t = (lane >> 1) & 7;
    const int my_pair = lane & 1;
    const int my_st0 = 4 * my_half + 2 * my_pair;                   // my stimuli: my_st0, my_st0 + 1
    const bool owner = my_t < RW_TI && (row0 + my_t < rows_here);
    const int grow = row_base + row0 + my_t;                        // global row of the owned outputs
    const int sid = lane & 7;                                       // stimulus whose status this lane tracks

    const unsigned x_local = smem_u32(Xf), xe_local = smem_u32(xe);
    const unsigned full_local[2] = {smem_u32(&misc->full[0]), smem_u32(&misc->full[1])};
    const unsigned xfull_local = smem_u32(&misc->xfull);
    // shared::cluster address of the same offset in CTA p = local address + pdelta[p]
    // (the cluster window of every CTA is laid out identically, so one subtraction gives the offset)
    if (tid < RW_MAXC) misc->pdelta[tid] = map_to_rank(x_local, (unsigned)(tid < csize ? tid : 0)) - x_local;
    const volatile unsigned *pdelta = misc->pdelta;

    // ---- one-time setup ----
    if (tid < 8) {
        double t = a.t_first;
        for (int l = 0; l < tid; ++l) t *= (1.0 / 64.0);
        misc->tlevel[tid] = (tid == 7 || t <= a.atol) ? 0.0 : t;
    }
    if (tid == 0) {
        mbar_init(full_local[0], 1 + RW_WARPS);          // the arming thread + one release-arrive per local warp
        mbar_init(full_local[1], 1 + RW_WARPS);
        mbar_init(xfull_local, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * RW_BUF_BYTES / 4; i += RW_THREADS) Xf[i] = 0.f;
    for (int i = tid; i < 2 * RW_XE * kpad; i += RW_THREADS) xe[i] = 0.f;
    if (a.w_kind == SSN_W_FROM_Z) build_profile_table(a.wc, N, gtab, tid, RW_THREADS);
    build_io_tables(a, tab, tid, RW_THREADS);
    cluster.sync();

    unsigned ph[2] = {0u, 0u}, xph = 0u;
    const unsigned tx_bytes = (a.dbg & 1) ? 0u : (unsigned)((csize - 1) * RW_WARPS * RW_BLK_BYTES);      // blocks arriving from peers
    // slot (16-byte unit inside a buffer) of panel column j = c*32 + lane, two per register
    unsigned colslot[(NC + 1) / 2];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int j = c * 32 + lane;
        unsigned slot = RW_ZERO_SLOT;
        if (j < dim) {
            const int cta = j / rpc, lr = j - cta * rpc, w = lr / RW_TI, t = lr - w * RW_TI;
            slot = (unsigned)((cta * RW_WARPS + w) * RW_BLK_SLOTS + t);
        }
        if (c & 1) colslot[c / 2] |= slot << 16; else colslot[c / 2] = slot;
    }
    const unsigned my_block = (unsigned)((rank * RW_WARPS + warp) * RW_BLK_BYTES);      // byte offset in a buffer
    const int n_chunks = (a.nb + TB - 1) / TB;

    for (;;) {
        // ---- next network from the global queue ----
        if (rank == 0 && tid == 0) {
            const int n = atomicAdd(a.work_counter, 1);
            for (int p = 0; p < csize; ++p) st_cluster_u32(map_to_rank(smem_u32(&misc->next_net), p), (unsigned)n);
        }
        cluster.sync();
        const int net = misc->next_net;
        if (net >= a.nz) break;

        // ---- W tile -> registers ----
        unsigned long long wp[NP][NC];                             // rows (0,1), (2,3), ... as packed pairs
        float ws[ODD ? NC : 1];                                    // the unpaired last row (odd TI)
        {
            // all 7 x NC loads are issued before any is consumed (one round trip to L2/HBM, not 98)
            const float *src = a.w + (size_t)net * dim * dim + (size_t)(row_base + row0) * dim + lane;
            float zv[RW_TI][NC];
#pragma unroll
            for (int t = 0; t < RW_TI; ++t)
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    zv[t][c] = (row0 + t < rows_here && c * 32 + lane < dim) ? __ldg(src + (size_t)t * dim + c * 32) : 0.f;
            if (a.w_kind == SSN_W_FROM_Z) {
#pragma unroll
                for (int t = 0; t < RW_TI; ++t) {
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
                for (int q = 0; q < NP; ++q) wp[q][c] = pack2(zv[2 * q][c], zv[2 * q + 1][c]);
                if (ODD) ws[c] = zv[TI - 1][c];
            }
        }

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int b0 = chunk * TB;
            const int nact = min(TB, a.nb - b0);
            const float *ext_net = a.ext + (size_t)net * a.ext_stride_z;

            // float64 state of the four (row, stimulus) outputs an owner lane holds lives in shared memory
            const double eps_own = grow < N ? a.eps_E : a.eps_I;
            unsigned levels = 0u;                            // refresh-ladder level of my two stimuli, 4 bits each
            float x0[2] = {0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double r0 = 0.0;
                float e = 0.f;
                const int st = my_st0 + i;
                if (owner && st < nact) {
                    e = __ldg(ext_net + (size_t)(b0 + st) * dim + grow);
                    if (a.r_init) r0 = (double)__ldg(a.r_init + ((size_t)net * a.nb + b0 + st) * dim + grow);
                }
                sR[i * RW_THREADS] = r0; sRref[i * RW_THREADS] = 0.0; sVref[i * RW_THREADS] = (double)e;
                sExt[i * RW_THREADS] = e;
                x0[i] = (float)r0;
            }
            const unsigned xoff = my_block + 16u * (unsigned)(my_t + RW_TI * my_half) + 8u * (unsigned)my_pair;   // my float2
            unsigned done = nact >= TB ? 0u : (0xffu << nact) & 0xffu;
            unsigned force_refresh = a.r_init ? (~done & 0xffu) : 0u;   // with r_init the first sweep refreshes
            int my_status = 1, my_iters = a.max_iter;                   // of stimulus `sid`

            // ---- publish the initial panel: r - r_ref (= r_init, refreshed at once, or 0) ----
            int buf = 0;
            if (tid == 0) mbar_arrive_expect_tx(full_local[0], tx_bytes);
            if (owner) *reinterpret_cast<float2 *>(smem + L.x_off + xoff) = make_float2(x0[0], x0[1]);
            if (lane == 0)
                *reinterpret_cast<unsigned *>(smem + L.x_off + my_block + 2 * RW_TI * 16) = 0x00ff0000u;  // "big": no refresh yet
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_release(full_local[0]);
            else if (lane <= csize - 1 && !(a.dbg & 1)) {
                const int p = lane - 1 + (lane - 1 >= rank ? 1 : 0);            // the csize-1 peers
                bulk_copy_to_peer(x_local + my_block + pdelta[p], x_local + my_block, RW_BLK_BYTES,
                                  full_local[0] + pdelta[p]);
            }

            long long tc[6] = {0, 0, 0, 0, 0, 0};
            for (int it = 1;; ++it) {
                long long c0 = clock64();
                // ---- wait for the panel of this sweep and the flags of the previous one ----
                mbar_wait(full_local[buf], ph[buf]);
                long long c1 = clock64(); tc[0] += c1 - c0;
                ph[buf] ^= 1u;
                unsigned F;
                {
                    const unsigned char *xb = smem + L.x_off + buf * RW_BUF_BYTES + 2 * RW_TI * 16;
                    unsigned f = *reinterpret_cast<const unsigned *>(xb + lane * RW_BLK_BYTES) |
                                 *reinterpret_cast<const unsigned *>(xb + (lane + 32) * RW_BLK_BYTES);
                    if (lane + 64 < RW_BLOCKS) f |= *reinterpret_cast<const unsigned *>(xb + (lane + 64) * RW_BLK_BYTES);
                    F = __reduce_or_sync(0xffffffffu, f);
                }
                if (a.dbg & 1) F = 0x00ff00ffu;          // timing experiment: no exchange, nobody converges or refreshes
                if (it > 1) {
                    const unsigned moving_all = F & 0xffu, above_all = (F >> 8) & 0xffu;
                    const unsigned conv_now = ~moving_all & ~done & 0xffu;       // ssnode.c:84-96 first ...
                    const unsigned hard_now = a.check_hard ? (above_all & ~done & ~conv_now & 0xffu) : 0u;  // ... then :98-102
                    if ((conv_now >> sid) & 1u) { my_status = 0; my_iters = it - 1; }
                    if ((hard_now >> sid) & 1u) { my_status = 2; my_iters = it - 1; }
                    done |= conv_now | hard_now;
                }
                if (done == 0xffu || it > a.max_iter) break;

                // ---- reference-point refresh for stimuli whose max|dr| fell below their ladder threshold ----
                const unsigned natural = ~(F >> 16) & ~done & 0xffu;
                unsigned req = natural | (force_refresh & ~done);
                {
                    // at most RW_XE stimuli per event; the rest keep asking and are served on the next sweeps
                    unsigned keep = 0u, left = req;
                    for (int q = 0; q < RW_XE && left; ++q) { keep |= left & (0u - left); left &= left - 1u; }
                    force_refresh &= ~keep;
                    req = keep;
                }
                if (req) {
                    const unsigned nreq = __popc(req);
                    if (tid == 0) mbar_arrive_expect_tx(xfull_local, nreq * (unsigned)dim * 8u);
                    if (owner) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int st = my_st0 + i;
                            if ((req >> st) & 1u) {
                                const double ri = sR[i * RW_THREADS];
                                const float hi = (float)ri;
                                const float lo = (float)(ri - (double)hi);
                                const unsigned o = 4u * (unsigned)(__popc(req & ((1u << st) - 1u)) * kpad + grow);
                                for (int p = 0; p < csize; ++p) {
                                    {
                                        const unsigned bar = xfull_local + pdelta[p];
                                        st_async_u32(xe_local + o + pdelta[p], __float_as_uint(hi), bar);
                                        st_async_u32(xe_local + o + 4u * (unsigned)(RW_XE * kpad) + pdelta[p],
                                                     __float_as_uint(lo), bar);
                                    }
                                }
                            }
                        }
                    }
                    mbar_wait(xfull_local, xph);
                    xph ^= 1u;
                    for (int s = 0; s < TB; ++s) {
                        if (!((req >> s) & 1u)) continue;
                        const int xs = __popc(req & ((1u << s) - 1u));
                        const float *xh = xe + xs * kpad, *xl = xe + (RW_XE + xs) * kpad;
                        double accd[RW_TI];
                        float accf[RW_TI];
#pragma unroll
                        for (int t = 0; t < RW_TI; ++t) { accd[t] = 0.0; accf[t] = 0.f; }
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            const double h = (double)xh[c * 32 + lane];
                            const float l = xl[c * 32 + lane];
                            float wv[RW_TI];
#pragma unroll
                            for (int q = 0; q < NP; ++q) unpack2(wp[q][c], wv[2 * q], wv[2 * q + 1]);
                            if (ODD) wv[TI - 1] = ws[c];
#pragma unroll
                            for (int t = 0; t < RW_TI; ++t) {
                                accd[t] = fma((double)wv[t], h, accd[t]);          // exact products, fp64 sum
                                accf[t] = fmaf(wv[t], l, accf[t]);
                            }
                        }
                        double mine = 0.0;
#pragma unroll
                        for (int t = 0; t < RW_TI; ++t) {
                            double v = accd[t] + (double)accf[t];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            mine = (t == my_t) ? v : mine;
                        }
                        if (owner && (s >> 1) == (my_st0 >> 1)) {
                            const int i = s & 1;
                            sVref[i * RW_THREADS] = mine + (double)sExt[i * RW_THREADS];
                            sRref[i * RW_THREADS] = sR[i * RW_THREADS];
                            if ((natural >> s) & 1u) levels += 1u << (4 * i);      // next rung of the ladder
                        }
                        // r - r_ref is now zero for this stimulus on every row of every CTA
                        float *col = reinterpret_cast<float *>(smem + L.x_off + buf * RW_BUF_BYTES) + (s & 3);
                        for (int q = tid; q < RW_BLOCKS * RW_TI; q += RW_THREADS) {
                            const int blk = q / RW_TI, t = q - blk * RW_TI;
                            col[4 * (blk * RW_BLK_SLOTS + t + RW_TI * (s >> 2))] = 0.f;
                        }
                    }
                    __syncthreads();
                }

                // ---- arm the next phase, then contract: dv = W * fl32(r - r_ref) ----
                const int nbuf = buf ^ 1;
                long long c2 = clock64(); tc[1] += c2 - c1;
                if (tid == 0) mbar_arrive_expect_tx(full_local[nbuf], tx_bytes);
                // Half-panel-major: the four stimuli of a half share one LDS.128 per column; a half whose stimuli
                // have all finished is skipped (uniform over the cluster).  Sixteen independent FMA chains per half.
                float acc[RW_TI][TB];
                {
                    const float4 *Xq = reinterpret_cast<const float4 *>(smem + L.x_off + buf * RW_BUF_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        unsigned long long ap[NP][4];
                        float as[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q = 0; q < NP; ++q) ap[q][b] = 0ull;
                            as[b] = 0.f;
                        }
                        if (((done >> (4 * h)) & 0xfu) != 0xfu) {
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                const unsigned slot = (c & 1) ? (colslot[c / 2] >> 16) : (colslot[c / 2] & 0xffffu);
                                const float4 x4 = Xq[slot + RW_TI * h];
                                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                                for (int b = 0; b < 4; ++b) {
#pragma unroll
                                    for (int q = 0; q < NP; ++q) ffma2(ap[q][b], wp[q][c], xv[b]);
                                    if (ODD) as[b] = fmaf(ws[c], xv[b], as[b]);
                                }
                            }
                        }
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q = 0; q < NP; ++q) unpack2(ap[q][b], acc[2 * q][4 * h + b], acc[2 * q + 1][4 * h + b]);
                            if (ODD) acc[TI - 1][4 * h + b] = as[b];
                        }
                    }
                }
                long long c3 = clock64(); tc[2] += c3 - c2;
                // ---- 32-lane reduce-scatter: stimulus half over lane bit 4, rows over bits 3,2,1, stimulus pair over bit 0 ----
                float dv[2];
                {
                    const unsigned full = 0xffffffffu;
                    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
                    float h8[8][4];                                    // rows 0..6 (+ a zero row), my half
#pragma unroll
                    for (int t = 0; t < RW_TI; ++t)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float send = u16 ? acc[t][c] : acc[t][4 + c];
                            const float keep = u16 ? acc[t][4 + c] : acc[t][c];
                            h8[t][c] = keep + __shfl_xor_sync(full, send, 16);
                        }
#pragma unroll
                    for (int t = RW_TI; t < 8; ++t)
#pragma unroll
                        for (int c = 0; c < 4; ++c) h8[t][c] = 0.f;
                    float h4[4][4], h2[2][4], h1[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float send = u8 ? h8[t][c] : h8[4 + t][c];
                            const float keep = u8 ? h8[4 + t][c] : h8[t][c];
                            h4[t][c] = keep + __shfl_xor_sync(full, send, 8);
                        }
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float send = u4 ? h4[t][c] : h4[2 + t][c];
                            const float keep = u4 ? h4[2 + t][c] : h4[t][c];
                            h2[t][c] = keep + __shfl_xor_sync(full, send, 4);
                        }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float send = u2 ? h2[0][c] : h2[1][c];
                        const float keep = u2 ? h2[1][c] : h2[0][c];
                        h1[c] = keep + __shfl_xor_sync(full, send, 2);
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float send = u1 ? h1[c] : h1[2 + c];
                        const float keep = u1 ? h1[2 + c] : h1[c];
                        dv[c] = keep + __shfl_xor_sync(full, send, 1);
                    }
                }
                // row index: bit 3 chose rows 4..7, bit 2 the upper pair, bit 1 the odd row  -> my_t = (lane >> 1) & 7

                long long c4 = clock64(); tc[3] += c4 - c3;
                // ---- float64 state update of the two outputs this lane owns ----
                unsigned word = 0u;
                if (owner) {
                    float xn[2];
                    // both outputs: loads and the branch-free common path first (two independent FP64 chains)
                    double vv[2], fv[2], r_old[2], r_ref2[2], tl[2];
                    bool rare[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        vv[i] = sVref[i * RW_THREADS] + (double)dv[i];
                        r_old[i] = sR[i * RW_THREADS];
                        r_ref2[i] = sRref[i * RW_THREADS];
                        tl[i] = misc->tlevel[(levels >> (4 * i)) & 7u];
                        fv[i] = io_eval_common(a, tab, vv[i], rare[i]);
                    }
                    if (rare[0] | rare[1]) {                                   // saturating / diverging neurons only
                        if (rare[0]) fv[0] = io_eval_exact(a, vv[0]);
                        if (rare[1]) fv[1] = io_eval_exact(a, vv[1]);
                    }
                    if (a.dbg & 2) { fv[0] = vv[0]; fv[1] = vv[1]; }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int st = my_st0 + i;
                        const bool live = !((done >> st) & 1u);
                        const double d = live ? (fv[i] - r_old[i]) * eps_own : 0.0;   // r_new - r_old
                        const double step = fabs(d);
                        const double r_cur = r_old[i] + d;
                        if (live && step >= a.atol) word |= 1u << st;
                        if (live && r_cur >= a.r_hard) word |= 1u << (8 + st);
                        if ((live && step >= tl[i]) || !(tl[i] > 0.0)) word |= 1u << (16 + st);   // exhausted ladder never asks
                        if (live) sR[i * RW_THREADS] = r_cur;
                        xn[i] = (float)(r_cur - r_ref2[i]);
                    }
                    *reinterpret_cast<float2 *>(smem + L.x_off + nbuf * RW_BUF_BYTES + xoff) = make_float2(xn[0], xn[1]);
                }
                long long c5 = clock64(); tc[4] += c5 - c4;
                word = __reduce_or_sync(0xffffffffu, word);
                if (lane == 0)
                    *reinterpret_cast<unsigned *>(smem + L.x_off + nbuf * RW_BUF_BYTES + my_block + 2 * RW_TI * 16) = word;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_release(full_local[nbuf]);
                else if (lane <= csize - 1 && !(a.dbg & 1)) {
                    const int p = lane - 1 + (lane - 1 >= rank ? 1 : 0);
                    const unsigned src = x_local + (unsigned)(nbuf * RW_BUF_BYTES) + my_block;
                    bulk_copy_to_peer(src + pdelta[p], src, RW_BLK_BYTES, full_local[nbuf] + pdelta[p]);
                }
                buf = nbuf;
                tc[5] += clock64() - c5;
            }
            if ((a.dbg & 4) && a.dbg_out && net == 0 && tid == 0)
                for (int q = 0; q < 6; ++q) a.dbg_out[rank * 6 + q] = tc[q];

            // ---- results ----
            if (owner) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int st = my_st0 + i;
                    if (st < nact) a.R[((size_t)net * a.nb + b0 + st) * dim + grow] = (float)sR[i * RW_THREADS];
                }
            }
            if (rank == 0 && warp == 0 && lane < nact) {               // lane == sid for lanes 0..7
                a.status[(size_t)net * a.nb + b0 + lane] = my_status;
                if (a.iters) a.iters[(size_t)net * a.nb + b0 + lane] = my_iters;
            }
            // nobody may publish the next panel while a slower CTA still reads this one
            cluster.sync();
        }
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
typedef void (*RwKernel)(const RwArgs);
static int rw_ti_for(int) { return 7; }                          // rows per warp of every instantiated shape
static RwKernel pick_rw_kernel(int nc, int nw) {
    if (nw == 4) {
        switch (nc) {
            case 2: return ssn_fp_regw_kernel<2, 4, 7>;
            case 4: return ssn_fp_regw_kernel<4, 4, 7>;
            case 7: return ssn_fp_regw_kernel<7, 4, 7>;
            case 10: return ssn_fp_regw_kernel<10, 4, 7>;
            case 14: return ssn_fp_regw_kernel<14, 4, 7>;
        }
        return nullptr;
    }
    // (14-warp CTAs with 6 rows per warp -- 5-CTA clusters at 2N = 402 -- need ~200 live registers against a
    //  budget of 144: ptxas spills 1.4 KB per thread, so that shape is not instantiated)
    switch (nc) {
        case 2: return ssn_fp_regw_kernel<2, 8, 7>;
        case 4: return ssn_fp_regw_kernel<4, 8, 7>;
        case 7: return ssn_fp_regw_kernel<7, 8, 7>;
        case 10: return ssn_fp_regw_kernel<10, 8, 7>;
        case 14: return ssn_fp_regw_kernel<14, 8, 7>;
    }
    return nullptr;
}

static int rw_nc_for(int dim) {
    const int cands[] = {2, 4, 7, 10, 14};
    for (int nc : cands)
        if (32 * nc >= dim) return nc;
    return 0;
}

struct RwPlan { RwKernel fn; int nc, nw, kpad, csize, rpc, smem, clusters, tab_nodes; };

static int plan_regw_nw(const ssn_solver &sv, int n_sites, int nz, int nw, RwPlan *plan) {
    const int dim = 2 * n_sites, ti = rw_ti_for(nw), rows = ti * nw;
    plan->nw = nw;
    plan->nc = rw_nc_for(dim);
    if (!plan->nc) return 1;                                   // too large: caller falls back to the smem kernel
    plan->kpad = 32 * plan->nc;
    plan->csize = (dim + rows - 1) / rows;
    if (plan->csize > (nw == 4 ? RW_MAXC : MAX_CLUSTER) || plan->csize * nw > RW_BLOCKS) return 1;
    plan->rpc = (dim + plan->csize - 1) / plan->csize;
    if (plan->rpc * (plan->csize - 1) >= dim) return 1;
    plan->tab_nodes = rw_table_nodes(sv);
    plan->smem = rw_smem_layout(plan->kpad, n_sites, rw_table_bytes(plan->tab_nodes, rw_table2_nodes(sv)), 32 * nw, ti).total;
    int dev = 0, limit = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (plan->smem > limit) return 1;
    plan->fn = pick_rw_kernel(plan->nc, nw);
    if (!plan->fn) return 1;
    SSN_CUDA(cudaFuncSetAttribute(plan->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem));
    if (plan->csize > 8) SSN_CUDA(cudaFuncSetAttribute(plan->fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan->csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(plan->csize, 1, 1);
    cfg.blockDim = dim3(32 * nw, 1, 1);
    cfg.dynamicSmemBytes = plan->smem;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, plan->fn, &cfg) != cudaSuccess) { cudaGetLastError(); return 1; }
    if (max_clusters < 1) return 1;
    plan->clusters = nz > 0 ? std::min(max_clusters, nz) : max_clusters;
    return 0;
}

// 8-warp CTAs in portable clusters (<= 8) by default; SSN_REGW_WARPS=4 selects 4-warp CTAs (two
// co-resident per SM, clusters of up to 16): measured equal in throughput at 2N = 402 on B200.
static int plan_regw(const ssn_solver &sv, int n_sites, int nz, RwPlan *plan) {
    const char *force = getenv("SSN_REGW_WARPS");
    const int first = force ? atoi(force) : 8;
    int rc = plan_regw_nw(sv, n_sites, nz, first == 4 ? 4 : 8, plan);
    if (rc == 1 && !force) rc = plan_regw_nw(sv, n_sites, nz, 4, plan);
    return rc;
}

int regw_occupancy(const ssn_solver &sv, int n_sites, int *cluster_size, int *resident_clusters) {
    RwPlan plan;
    int rc = plan_regw(sv, n_sites, 0, &plan);
    if (rc) return rc;
    if (cluster_size) *cluster_size = plan.csize;
    if (resident_clusters) *resident_clusters = plan.clusters;
    return 0;
}

// Returns 1 when the shape is outside this kernel's range (the caller then uses the
// shared-memory kernel), 0 on success, otherwise an error code.
int launch_fixed_point_regw(const ssn_solver &sv, int nz, int nb, int n_sites, int w_kind, const float *w,
                            const ssn_jds *jds, const float *ext, int ext_per_network, const float *r_init,
                            float *R, int *status, int *iters, int *counter, cudaStream_t stream) {
    RwPlan plan;
    int rc = plan_regw(sv, n_sites, nz, &plan);
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
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(plan.clusters * plan.csize, 1, 1);
    cfg.blockDim = dim3(32 * plan.nw, 1, 1);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (a.dbg & 4) SSN_CUDA(cudaMalloc(&a.dbg_out, 128 * sizeof(long long)));
    SSN_CUDA(cudaLaunchKernelEx(&cfg, plan.fn, a));
    count_launch();
    if (a.dbg & 4) {
        long long h[48];
        SSN_CUDA(cudaStreamSynchronize(stream));
        SSN_CUDA(cudaMemcpy(h, a.dbg_out, sizeof(h), cudaMemcpyDeviceToHost));
        const char *names[6] = {"wait", "top/refresh", "contract", "reduce", "update", "publish"};
        fprintf(stderr, "[ssn dbg] warps/CTA %d cluster %d resident clusters %d smem %d\n", plan.nw, plan.csize, plan.clusters, plan.smem);
        for (int r = 0; r < (plan.csize < 8 ? plan.csize : 8); r += 7) {
            fprintf(stderr, "[ssn dbg] rank %d cycles (net 0, all sweeps):", r);
            for (int q = 0; q < 6; ++q) fprintf(stderr, " %s=%lld", names[q], h[r * 6 + q]);
            fprintf(stderr, "\n");
        }
        cudaFree(a.dbg_out);
    }
    return 0;
}

}  // namespace ssn
