// Pieces of the register-resident-W fixed-point kernel (ssn_fixed_point_ws.cu): kernel arguments, mbarrier /
// st.async / bulk-copy wrappers, packed-pair FMA, and the float64 table evaluation of the transfer function.
#pragma once
#include <cmath>
#include <cstdlib>
#include "ssn_cluster_core.cuh"

namespace ssn {

constexpr int TAB_PER_UNIT = 8;                        // nodes per unit of v of the power-law table
constexpr double TAB_V_MIN = 1.0;

struct RwArgs {
    int nz, nb, n_sites, dim, kpad, csize, rpc;
    int w_kind;
    const float *w;
    WeightConst wc;
    const float *ext;
    long long ext_stride_z;
    const float *r_init;
    float *R;
    int *status, *iters;
    int *work_counter;
    IoConst<double> io;
    IoConst<float> iof;
    double eps_E, eps_I, atol, r_hard, t_first;        // t_first: first refresh threshold on |dr|
    int max_iter, check_hard, tab_nodes, tab2_nodes;   // nodes of the power-law table / of the tanh table (0: none)
    float tab_end;                                     // v at the end of the power-law table
    double tab2_end;                                   // v at the end of the tanh table
    int dbg;                                           // development switches (SSN_DBG), 0 in production
    long long *dbg_out;                                // phase cycle counters of the profile builds
};

// ---- mbarrier / st.async helpers ---------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef SSN_MBAR_HINT_NS
#define SSN_MBAR_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
#if SSN_MBAR_HINT_NS > 0
        // suspend-time hint: the warp may stay suspended this long before try_wait returns false (it is woken when the
        // phase completes), so a waiting warp polls -- and takes issue slots from its scheduler -- less often
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;"
            " selp.u32 %0, 1, 0, p; }"
            : "=r"(done) : "r"(bar), "r"(parity), "r"((unsigned)SSN_MBAR_HINT_NS) : "memory");
#else
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;"
            " selp.u32 %0, 1, 0, p; }"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#endif
    }
}
__device__ __forceinline__ void st_async_v4(unsigned addr, float4 v, unsigned bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                   "r"(__float_as_uint(v.w)), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_v2(unsigned addr, float a, float b, unsigned bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release(unsigned bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_copy_to_peer(unsigned dst_remote, unsigned src_local, unsigned bytes, unsigned bar_remote) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_remote), "r"(src_local), "r"(bytes), "r"(bar_remote) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void st_async_u32(unsigned addr, unsigned v, unsigned bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(addr), "r"(v), "r"(bar) : "memory");
}

// packed FP32 pairs (FFMA2): a register pair holds rows (2p, 2p+1) of the W tile / of the accumulators, so one
// fma.rn.f32x2 with the broadcast x does two FMAs and its operands can never collide on a register bank
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long w, float x) {
    unsigned long long xx;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(xx));
}

// ---- f(v) in float64 from cubic Taylor tables in shared memory ------------------------------------
// Table 1: k v^n on [1, v_end], TAB_PER_UNIT nodes per unit of v.  Table 2 (asym_tanh only): the saturating
// branch r0 + (r1 - r0) tanh(.) on [v0, v0 + (TAB2_NODES - 1) / TAB2_PER_UNIT].  A node holds the four Taylor
// coefficients c_j = f^(j)(node) h^j / j!; truncation error < 3e-9 (table 1, at v = 1; 1e-10 and less from
// v = 2 on) and < 1e-9 (table 2), absolute, in rate units.  Node index by the 2^52 rounding trick (two DADD, no
// conversion instructions), Horner in FP64, no branches: saturated neurons (rates above rate_soft_bound) cost
// the same as the others.  v < 1: FP32 power (|f| < k, absolute error ~1e-8 k).  Beyond the tables (diverging
// networks, rates within 0.5 % of the hard bound): closed form behind a rare branch (io_eval_exact).
constexpr int TAB_COEF = 4;
constexpr int TAB2_PER_UNIT = 2;
constexpr int TAB2_NODES = 1025;
__host__ __device__ constexpr int rw_table_bytes(int nodes1, int nodes2) { return (nodes1 + nodes2) * TAB_COEF * 8; }

template <class Args>
__device__ __forceinline__ void build_io_tables(const Args &a, double *tab, int tid, int nthreads) {
    for (int i = tid; i < a.tab_nodes; i += nthreads) {
        const double v = TAB_V_MIN + (double)i / TAB_PER_UNIT, h = 1.0 / TAB_PER_UNIT;
        const double n = a.io.n, p3 = pow(v, n - 3.0);
        tab[4 * i + 0] = a.io.k * p3 * v * v * v;
        tab[4 * i + 1] = a.io.k * n * p3 * v * v * h;
        tab[4 * i + 2] = a.io.k * n * (n - 1.0) * p3 * v * h * h * 0.5;
        tab[4 * i + 3] = a.io.k * n * (n - 1.0) * (n - 2.0) * p3 * h * h * h / 6.0;
    }
    for (int i = tid; i < a.tab2_nodes; i += nthreads) {
        // derivatives of T = tanh(y) with S = 1 - T^2:  S,  -2 T S,  -2 S (1 - 3 T^2)
        const double h = 1.0 / TAB2_PER_UNIT, y = a.io.tanh_scale * h * i, T = tanh(y), S = 1.0 - T * T, g = a.io.tanh_scale * h;
        double *c = tab + TAB_COEF * (a.tab_nodes + i);
        c[0] = a.io.r_soft + a.io.span * T;
        c[1] = a.io.span * g * S;
        c[2] = a.io.span * g * g * (-T * S);
        c[3] = a.io.span * g * g * g * (-S * (1.0 - 3.0 * T * T)) / 3.0;
    }
}

// closed form, for the rare values outside the tables
template <class Args>
__device__ __noinline__ double io_eval_exact(const Args &a, double v) {
    if (!(v > 0.0)) return v != v ? v : 0.0;
    if (a.io.io_type != SSN_IO_POWER && v > a.io.v0)
        return a.io.io_type == SSN_IO_LINEAR ? fma(a.io.lin_slope, v - a.io.v0, a.io.r_soft)
                                             : a.io.r_soft + a.io.span * tanh(a.io.tanh_scale * (v - a.io.v0));
    return a.io.k * pow(v, a.io.n);
}

// table evaluation; `rare` reports that v lies outside the tables and io_eval_exact must be used instead
template <class Args>
__device__ __forceinline__ double io_eval_common(const Args &a, const double *tab, double v, bool &rare) {
    const float vf = (float)v;
    const bool above = v > a.io.v0;
    const bool upper = above && a.tab2_nodes > 0;                       // saturating branch of asym_tanh
    const double inv_h = upper ? (double)TAB2_PER_UNIT : (double)TAB_PER_UNIT;
#ifndef SSN_WS_FEVAL
#define SSN_WS_FEVAL 0
#endif
#if SSN_WS_FEVAL == 1
    // Experiment (not the default): node index from FP32 arithmetic + F2I / I2D conversions, s in two dependent FP64
    // operations instead of five, Estrin form.  Measured 2.8 % SLOWER at configs[1] (75.64 against 73.59 ms): the
    // conversion instructions cost more than the three FP64 operations they replace (the 2^52 trick below has none);
    // Estrin alone (SSN_WS_FEVAL=2) changes nothing (73.78 against 73.71 ms).
    int i = __float2int_rn((vf - (upper ? a.iof.v0 : (float)TAB_V_MIN)) * (upper ? (float)TAB2_PER_UNIT : (float)TAB_PER_UNIT));
    i = max(0, min(i, (upper ? a.tab2_nodes : a.tab_nodes) - 1));
    const double *c = tab + TAB_COEF * (i + (upper ? a.tab_nodes : 0));
    const double2 c01 = *reinterpret_cast<const double2 *>(c);
    const double2 c23 = *reinterpret_cast<const double2 *>(c + 2);
    const double s = fma(v - (upper ? a.io.v0 : TAB_V_MIN), inv_h, -(double)i);
    double f = fma(s * s, fma(s, c23.y, c23.x), fma(s, c01.y, c01.x));
#else
    const double x = (v - (upper ? a.io.v0 : TAB_V_MIN)) * inv_h;
    // round to nearest by adding 1.5 * 2^52: the integer lands in the low word, the rounded value comes back by subtraction
    const double xm = x + 6755399441055744.0;
    int i = __double2loint(xm);
    const double s = x - (xm - 6755399441055744.0);
    i = max(0, min(i, (upper ? a.tab2_nodes : a.tab_nodes) - 1));
    const double *c = tab + TAB_COEF * (i + (upper ? a.tab_nodes : 0));
    const double2 c01 = *reinterpret_cast<const double2 *>(c);
    const double2 c23 = *reinterpret_cast<const double2 *>(c + 2);
#if SSN_WS_FEVAL == 2
    double f = fma(s * s, fma(s, c23.y, c23.x), fma(s, c01.y, c01.x));      // Estrin: depth 2
#else
    double f = fma(s, fma(s, fma(s, c23.y, c23.x), c01.y), c01.x);
#endif
#endif
    const float flow = a.iof.k * exp2f(a.iof.n * __log2f(fmaxf(vf, 1e-30f)));
    f = vf < (float)TAB_V_MIN ? (double)flow : f;
    f = v > 0.0 ? f : (v != v ? v : 0.0);
    const bool lin_upper = above && a.io.io_type == SSN_IO_LINEAR;
    f = lin_upper ? fma(a.io.lin_slope, v - a.io.v0, a.io.r_soft) : f;
    rare = upper ? v >= a.tab2_end : (vf >= a.tab_end && !lin_upper);
    return f;
}

// Host: the solver-dependent kernel arguments (everything but the shape / plan fields).
inline int rw_table_nodes(const ssn_solver &sv) {
    // table of k v^n on [1, min(v0, 160)] (power type: to 160, beyond it the closed form is used)
    const double v0 = pow(sv.rate_soft_bound / sv.k, 1.0 / sv.n);
    double v_end = (sv.io_type == SSN_IO_POWER || !(v0 < 160.0)) ? 160.0 : v0 + 1.0;
    if (!(v_end > 2.0)) v_end = 2.0;
    return (int)((v_end - TAB_V_MIN) * TAB_PER_UNIT) + 2;
}
inline int rw_table2_nodes(const ssn_solver &sv) { return sv.io_type == SSN_IO_TANH ? TAB2_NODES : 0; }
inline void rw_fill_solver_args(RwArgs &a, const ssn_solver &sv, int tab_nodes) {
    a.io = make_io_const<double>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.iof = make_io_const<float>(sv.io_type, sv.k, sv.n, sv.rate_soft_bound, sv.rate_hard_bound);
    a.eps_E = sv.dt / sv.tau_E; a.eps_I = sv.dt / sv.tau_I;
    a.atol = sv.atol; a.r_hard = sv.rate_hard_bound;
    a.max_iter = sv.max_iter; a.check_hard = sv.io_type != SSN_IO_TANH;
    a.tab_nodes = tab_nodes;
    a.tab_end = (float)(TAB_V_MIN + (double)(tab_nodes - 1) / TAB_PER_UNIT - 0.5 / TAB_PER_UNIT);
    a.tab2_nodes = rw_table2_nodes(sv);
    a.tab2_end = a.io.v0 + (TAB2_NODES - 1.5) / TAB2_PER_UNIT;
    a.dbg = getenv("SSN_DBG") ? atoi(getenv("SSN_DBG")) : 0;
    // refresh ladder: thresholds atol * 64^j, starting at the largest one below 0.1
    double t = sv.atol > 0 ? sv.atol : 1e-300;
    while (t * 64.0 < 0.1) t *= 64.0;
    a.t_first = t > sv.atol ? t : 0.0;
}

}  // namespace ssn
