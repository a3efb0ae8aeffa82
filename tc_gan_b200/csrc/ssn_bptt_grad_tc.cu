// K4b on the 5th-generation tensor cores: the parameter gradient of the unrolled Euler SSN.
//
//   dL/dW = sum_k adj_k traj_k^T   (per network a [2N x K] x [K x 2N] contraction, K = seqlen * nb = 9600 at
//                                   BASELINE configs[2]),   dL/dtheta = < dL/dW, dW/dtheta >,  theta = J, D, S
//
// This is the one dense contraction of the SSN path (what Theano's autodiff through the scan of
// tc_gan/networks/ssn.py:354-385, 555-576 produces for J, D, S via make_w_batch.py:8-121).  The FP32 FFMA
// version (ssn_bptt_param_grad_kernel) ran at 29 TFLOP/s = 39 % of the FFMA roofline; here:
//
//   * operands come straight from the BPTT scratch arrays by TMA (cp.async.bulk.tensor, 128-byte swizzle with
//     32-byte atoms: the only swizzled MN-major layout 32-bit operands may use, SWIZZLE_128B_BASE32B).
//     Both are "MN-major" for the MMA: adj [K][pitch] has the M index (row of W) contiguous, traj [K][pitch]
//     the N index (column of W) contiguous, so a TMA box {32 floats of M or N, BK rows of K} IS the canonical
//     MN-major operand layout: one box per 32-wide block of M / N, blocks LBO = BK * 128 bytes apart, the 4-row
//     swizzle groups along K SBO = 512 bytes apart;
//   * tcgen05.mma kind::tf32 with the 3-term split  a b ~ ah bh + ah bl + al bh  (ah = a with the low 13
//     mantissa bits cleared, al = a - ah exactly): a warp group rewrites each landed stage in place as `hi` and
//     writes `lo` to a twin buffer (an elementwise map, so the swizzle never has to be undone), which keeps
//     FP32-level accuracy (2^-21 per product) on the tensor pipe;
//   * accumulators (128 x 208 fp32) live in TMEM, two of them, so the epilogue of one output tile overlaps
//     the MMAs of the next; the epilogue reads them with tcgen05.ld and contracts on the fly against
//     dW/dJ, dW/dD, dW/dS (z re-read, Gaussian profile from a table): dL/dW is never written;
//   * K is cut into chunks of CHUNK_STEPS * BK = 2048 rows, each with its own accumulator and epilogue (the
//     contraction against dW/dtheta is linear, so chunks simply add up in the float64 result): the tensor core's
//     FP32 accumulation is not round-to-nearest, and over all 9600 rows its bias reached 5e-5 of the result;
//   * persistent CTAs (one per SM) walk the (network, K chunk, 128 x 208 tile) list; warp roles: 0 = TMA producer,
//     1 = MMA issuer (one elected lane), 4..11 = hi/lo splitters, 12..15 = epilogue.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include "ssn_common.cuh"
#include "ssn_launch.h"

namespace ssn {

namespace tc {

constexpr int BM = 128, BN = 208, BK = 16;             // output tile and K step
constexpr int BOXW = 32;                                // floats per TMA box row = one 128-byte swizzle row
constexpr int A_BOXES = BM / BOXW;                      // 4
constexpr int B_BOXES = (BN + BOXW - 1) / BOXW;         // 7 (224 columns loaded, 208 used)
constexpr int BOX_BYTES = BOXW * BK * 4;                // 2048
constexpr int A_BYTES = A_BOXES * BOX_BYTES;            // 8192
constexpr int B_BYTES = B_BOXES * BOX_BYTES;            // 14336
constexpr int HALF_BYTES = A_BYTES + B_BYTES;           // 22528: the hi (= TMA destination) or the lo operands
constexpr int STAGE_BYTES = 2 * HALF_BYTES;             // 45056
constexpr int STAGES = 4;
constexpr int CHUNK_STEPS = 128;                        // K steps accumulated in TMEM before an epilogue
constexpr int THREADS = 512;
constexpr int SPLIT_WARP0 = 4, SPLIT_WARPS = 8, EPI_WARP0 = 12, EPI_WARPS = 4;
constexpr int TMEM_COLS = 512, ACC_COLS = 256;          // two accumulators of BN (<= 256) columns
constexpr int UMMA_K = 8;                               // K of one kind::tf32 instruction
constexpr unsigned SBO = 512, LBO = BK * 128;           // stride of the 4-row swizzle groups along K, of the MN blocks (bytes)
constexpr unsigned KSTEP_BYTES = UMMA_K * 128;          // one MMA consumes 8 K rows of 128 bytes

struct Barriers {
    unsigned long long full[STAGES];        // TMA landed (tx bytes)
    unsigned long long ready[STAGES];       // hi/lo written and fenced (one arrival per splitter warp)
    unsigned long long empty[STAGES];       // MMAs that read the stage have completed (tcgen05.commit)
    unsigned long long acc_full[2];         // accumulator complete (tcgen05.commit)
    unsigned long long acc_empty[2];        // accumulator drained (one arrival per epilogue warp)
    unsigned tmem_base;
};

constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 1024 /* barriers + tables */ + 4 * 4 * 256;

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol error (wrong byte count, lost arrival) becomes a trap and a CUDA error after ~2 s
// instead of a hung device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    long long t0 = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;"
            " selp.u32 %0, 1, 0, p; }"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && (spins & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) asm volatile("trap;");          // ~2 s: a lost arrival, not a slow one
        }
    }
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int x, int y, int z, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// shared-memory matrix descriptor: MN-major, SWIZZLE_128B_BASE32B (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3fffu);
    d |= (unsigned long long)((LBO >> 4) & 0x3fffu) << 16;
    d |= (unsigned long long)((SBO >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;                                    // descriptor version (Blackwell)
    d |= 1ull << 61;                                    // SWIZZLE_128B_BASE32B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both MN-major, M x N
constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                          unsigned accumulate) {
    asm volatile(
        "{ .reg .pred p; setp.ne.b32 p, %4, 0;"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Params {
    int nz, n_sites, dim, tiles_m, tiles_n;
    int k_steps, k_chunks;              // K steps of BK rows in all, chunks of <= CHUNK_STEPS steps
    long long K;
    const float *z;
    WeightConst wc;
    double *grad;
    float *dbg;                         // development: raw operands / accumulators of the first tile (SSN_K4B_DEBUG)
};

__global__ void __launch_bounds__(THREADS, 1)
ssn_bptt_param_grad_tc_kernel(const __grid_constant__ CUtensorMap map_adj, const __grid_constant__ CUtensorMap map_traj,
                              const Params p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment: the 128-byte swizzle is a function of the absolute shared-memory address
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    unsigned char *stages = smem;
    Barriers *bars = reinterpret_cast<Barriers *>(smem + STAGES * STAGE_BYTES);
    float *gtab = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES + 512);          // [4][<= 256] Gaussian profile
    __shared__ double red[12];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dim = p.dim, N = p.n_sites;
    // work item = (network, K chunk, output tile); items of one (network, chunk) are adjacent, so CTAs that run
    // side by side read the same rows of adj / traj from L2
    const int tiles_per_net = p.tiles_m * p.tiles_n;
    const int items_per_net = tiles_per_net * p.k_chunks;
    const int n_tiles = p.nz * items_per_net;
    struct Item { int net, i0, j0, ks0, ks1; };
    auto item_of = [&](int w) {
        Item it;
        it.net = w / items_per_net;
        const int r = w - it.net * items_per_net, ch = r / tiles_per_net, tt = r - ch * tiles_per_net;
        it.i0 = (tt / p.tiles_n) * BM; it.j0 = (tt % p.tiles_n) * BN;
        it.ks0 = ch * CHUNK_STEPS; it.ks1 = min(p.k_steps, it.ks0 + CHUNK_STEPS);
        return it;
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_addr(&bars->full[s]), 1);
            mbar_init(smem_addr(&bars->ready[s]), SPLIT_WARPS);
            mbar_init(smem_addr(&bars->empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_addr(&bars->acc_full[a]), 1);
            mbar_init(smem_addr(&bars->acc_empty[a]), EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 12) red[tid] = 0.0;
    if (N <= 256) build_profile_table(p.wc, N, gtab, tid, THREADS);
    if (warp == 1) {                                    // TMEM: this warp allocates and (at the end) frees
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_addr(&bars->tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            unsigned it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const Item w = item_of(tile);
                const int net = w.net, i0 = w.i0, j0 = w.j0;
                for (int ks = w.ks0; ks < w.ks1; ++ks, ++it) {
                    const int s = it % STAGES;
                    const unsigned ph = (it / STAGES) & 1u;
                    mbar_wait(smem_addr(&bars->empty[s]), ph ^ 1u);            // slot free (first pass: immediately)
                    const unsigned bar = smem_addr(&bars->full[s]);
                    mbar_expect_tx(bar, HALF_BYTES);
                    const unsigned dst = smem_addr(stages + s * STAGE_BYTES);
#pragma unroll
                    for (int b = 0; b < A_BOXES; ++b)
                        tma_load_3d(dst + b * BOX_BYTES, &map_adj, i0 + b * BOXW, ks * BK, net, bar);
#pragma unroll
                    for (int b = 0; b < B_BOXES; ++b)
                        tma_load_3d(dst + A_BYTES + b * BOX_BYTES, &map_traj, j0 + b * BOXW, ks * BK, net, bar);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            unsigned it = 0, local_tile = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local_tile) {
                const unsigned acc = local_tile & 1u, acc_ph = (local_tile >> 1) & 1u;
                mbar_wait(smem_addr(&bars->acc_empty[acc]), acc_ph ^ 1u);      // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d_tmem = tmem_base + acc * ACC_COLS;
                const Item w = item_of(tile);
                for (int ks = w.ks0; ks < w.ks1; ++ks, ++it) {
                    const int s = it % STAGES;
                    const unsigned ph = (it / STAGES) & 1u;
                    mbar_wait(smem_addr(&bars->ready[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned a_hi = smem_addr(stages + s * STAGE_BYTES), b_hi = a_hi + A_BYTES;
                    const unsigned a_lo = a_hi + HALF_BYTES, b_lo = b_hi + HALF_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                        const unsigned o = kk * KSTEP_BYTES;                    // next group of 8 K rows
                        umma_tf32(d_tmem, umma_desc(a_hi + o), umma_desc(b_hi + o), ((ks - w.ks0) | kk) != 0);
                        umma_tf32(d_tmem, umma_desc(a_hi + o), umma_desc(b_lo + o), 1u);
                        umma_tf32(d_tmem, umma_desc(a_lo + o), umma_desc(b_hi + o), 1u);
                    }
                    umma_commit(smem_addr(&bars->empty[s]));                    // frees the stage when these MMAs are done
                }
                umma_commit(smem_addr(&bars->acc_full[acc]));
            }
        }
    } else if (warp >= SPLIT_WARP0 && warp < SPLIT_WARP0 + SPLIT_WARPS) {
        // ===== splitters: hi = x with the low 13 mantissa bits cleared (in place), lo = x - hi (twin buffer) =====
        const int st = tid - SPLIT_WARP0 * 32;
        unsigned it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const Item w = item_of(tile);
            for (int ks = w.ks0; ks < w.ks1; ++ks, ++it) {
                const int s = it % STAGES;
                const unsigned ph = (it / STAGES) & 1u;
                mbar_wait(smem_addr(&bars->full[s]), ph);
                float4 *hi = reinterpret_cast<float4 *>(stages + s * STAGE_BYTES);
                float4 *lo = reinterpret_cast<float4 *>(stages + s * STAGE_BYTES + HALF_BYTES);
                if (p.dbg && blockIdx.x == 0 && it == 0 && st < 256) {          // first stage as landed: A box 0, B box 0
                    p.dbg[st] = reinterpret_cast<const float *>(hi)[st];
                    p.dbg[256 + st] = reinterpret_cast<const float *>(hi)[A_BYTES / 4 + st];
                }
#pragma unroll 2
                for (int c = st; c < HALF_BYTES / 16; c += SPLIT_WARPS * 32) {
                    const float4 x = hi[c];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
                    hi[c] = h;
                    lo[c] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_addr(&bars->ready[s]));
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue: < dL/dW tile, dW/dtheta > straight from TMEM =====
        const int q = warp & 3;                             // TMEM lane quadrant this warp may read
        unsigned local_tile = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local_tile) {
            const unsigned acc = local_tile & 1u, acc_ph = (local_tile >> 1) & 1u;
            const Item w = item_of(tile);
            const int net = w.net, i0 = w.i0, j0 = w.j0;
            mbar_wait(smem_addr(&bars->acc_full[acc]), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int i = i0 + q * 32 + lane;                // row of W handled by this thread
            const bool row_ok = i < dim;
            const int ah = i >= N, ii = i - ah * N;
            const float *z_row = p.z + ((size_t)net * dim + (row_ok ? i : 0)) * dim;
            float sJ[2] = {0.f, 0.f}, sD[2] = {0.f, 0.f}, sS[2] = {0.f, 0.f};       // by column half bh
            const unsigned taddr = tmem_base + acc * ACC_COLS + ((unsigned)(q * 32) << 16);
            for (int c0 = 0; c0 < BN; c0 += 16) {
                // z of these 16 columns first (independent of the accumulator): their latency overlaps the TMEM load
                float zz16[16];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const int j = j0 + c0 + e;                   // rows of z are 8-byte aligned (2N even), j even
                    float2 zp = make_float2(0.f, 0.f);
                    if (row_ok && j + 1 < dim) zp = __ldg(reinterpret_cast<const float2 *>(z_row + j));
                    else if (row_ok && j < dim) zp.x = __ldg(z_row + j);
                    zz16[e] = zp.x; zz16[e + 1] = zp.y;
                }
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (p.dbg && tile == 0 && c0 < 32)
                    for (int e = 0; e < 16; ++e) p.dbg[512 + (q * 32 + lane) * 32 + c0 + e] = v[e];
                if (row_ok) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int j = j0 + c0 + e;
                        if (j < dim) {
                            const int bh = j >= N, ab = ah * 2 + bh;
                            int d = ii - (j - bh * N);
                            d = d < 0 ? -d : d;
                            const float x = (float)d * p.wc.dx;
                            const float g = N <= 256 ? gtab[ab * N + d] : expf(-x * x * p.wc.inv2s2[ab]);
                            const float zz = zz16[e];
                            const float gG = g * v[e];
                            // signs and 1/S^3 are applied once per (a, b) block below
                            sJ[bh] += gG;
                            sD[bh] = fmaf(gG, zz, sD[bh]);
                            sS[bh] = fmaf(gG * x * x, fmaf(p.wc.sD[ab], zz, p.wc.sJ[ab]), sS[bh]);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(&bars->acc_empty[acc]));
            // all rows of a warp lie in one row half only if the quadrant does not straddle N: reduce per (ah, bh)
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int bh = 0; bh < 2; ++bh) {
                    const bool mine = row_ok && ah == a2;
                    float vJ = mine ? sJ[bh] : 0.f, vD = mine ? sD[bh] : 0.f, vS = mine ? sS[bh] : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        vJ += __shfl_xor_sync(0xffffffffu, vJ, o);
                        vD += __shfl_xor_sync(0xffffffffu, vD, o);
                        vS += __shfl_xor_sync(0xffffffffu, vS, o);
                    }
                    if (lane == 0 && (vJ != 0.f || vD != 0.f || vS != 0.f)) {
                        const int ab = a2 * 2 + bh;
                        const float sgn = bh == 0 ? 1.f : -1.f;
                        atomicAdd(&red[ab], (double)(sgn * vJ));
                        atomicAdd(&red[4 + ab], (double)(sgn * vD));
                        atomicAdd(&red[8 + ab], (double)(vS * p.wc.invS3[ab]));
                    }
                }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 12 && red[tid] != 0.0) atomicAdd(p.grad + tid, red[tid]);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled encode_fn() {
    static EncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiled)p;
        else
            cudaGetLastError();
    });
    return fn;
}

// [nz][K][pitch] float32, inner extent `dim` (columns dim..pitch-1 and everything out of range read as zero)
static int make_map(CUtensorMap *map, const float *base, int nz, long long K, int dim, int pitch) {
    EncodeTiled enc = encode_fn();
    if (!enc) return 1;
    const cuuint64_t gdim[3] = {(cuuint64_t)dim, (cuuint64_t)K, (cuuint64_t)nz};
    const cuuint64_t gstride[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)K * pitch * 4};
    const cuuint32_t box[3] = {BOXW, BK, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, gdim, gstride, box, estride,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (base %p, dim %d, K %lld, pitch %d)", (int)r, (const void *)base,
                  dim, K, pitch);
        return -1;
    }
    return 0;
}

}  // namespace tc

int launch_bptt_param_grad_tc(int nz, int n_sites, long long K, int pitch, const float *adj, const float *traj,
                              const float *z, const WeightConst &wc, double *grad, cudaStream_t stream) {
    using namespace tc;
    const int dim = 2 * n_sites;
    if ((pitch & 3) || ((size_t)adj & 15) || ((size_t)traj & 15)) return 1;      // TMA needs 16-byte rows
    CUtensorMap map_adj, map_traj;
    int rc = make_map(&map_adj, adj, nz, K, dim, pitch);
    if (rc) return rc;
    if ((rc = make_map(&map_traj, traj, nz, K, dim, pitch))) return rc;
    Params p = {};
    p.nz = nz; p.n_sites = n_sites; p.dim = dim; p.K = K;
    p.tiles_m = (dim + BM - 1) / BM; p.tiles_n = (dim + BN - 1) / BN;
    p.k_steps = (int)((K + BK - 1) / BK);
    p.k_chunks = (p.k_steps + CHUNK_STEPS - 1) / CHUNK_STEPS;
    p.z = z; p.wc = wc; p.grad = grad;
    static float *dbg_buf = nullptr;
    if (getenv("SSN_K4B_DEBUG")) {
        if (!dbg_buf) SSN_CUDA(cudaMalloc(&dbg_buf, (512 + 128 * 32) * sizeof(float)));
        SSN_CUDA(cudaMemsetAsync(dbg_buf, 0, (512 + 128 * 32) * sizeof(float), stream));
        p.dbg = dbg_buf;
    }
    int dev = 0, sms = 0;
    SSN_CUDA(cudaGetDevice(&dev));
    SSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SSN_CUDA(cudaFuncSetAttribute(ssn_bptt_param_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const long long n_items = (long long)nz * p.tiles_m * p.tiles_n * p.k_chunks;
    if (n_items > 0x7fffffffll) { set_error("parameter-gradient contraction: too many work items"); return -1; }
    const int grid = n_items < sms ? (int)n_items : sms;
    {
        KernelTimer kt("ssn_bptt_param_grad_tc_kernel", stream);
        ssn_bptt_param_grad_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(map_adj, map_traj, p);
        SSN_CUDA(cudaGetLastError());
    }
    count_launch();
    if (p.dbg) {
        static float h[512 + 128 * 32];
        SSN_CUDA(cudaStreamSynchronize(stream));
        SSN_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[k4b dbg] A box0 row0:");
        for (int i = 0; i < 8; ++i) fprintf(stderr, " %.4f", h[i]);
        fprintf(stderr, "\n[k4b dbg] A box0 row1:");
        for (int i = 0; i < 8; ++i) fprintf(stderr, " %.4f", h[32 + i]);
        fprintf(stderr, "\n[k4b dbg] B box0 row0:");
        for (int i = 0; i < 8; ++i) fprintf(stderr, " %.4f", h[256 + i]);
        fprintf(stderr, "\n[k4b dbg] acc row0:");
        for (int i = 0; i < 8; ++i) fprintf(stderr, " %.4f", h[512 + i]);
        fprintf(stderr, "\n[k4b dbg] acc row1:");
        for (int i = 0; i < 8; ++i) fprintf(stderr, " %.4f", h[512 + 32 + i]);
        fprintf(stderr, "\n[k4b dbg] acc row5 col 16..:");
        for (int i = 0; i < 8; ++i) fprintf(stderr, " %.4f", h[512 + 5 * 32 + 16 + i]);
        fprintf(stderr, "\n");
    }
    return 0;
}

}  // namespace ssn
