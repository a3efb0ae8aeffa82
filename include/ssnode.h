/*
 * ssnode.h -- C ABI of the B200-native SSN library (libssnode.so).
 *
 * The library is a drop-in for the reference C extension tc_gan/ext/libssnode.so
 * (built from tc_gan/ext/ssnode.c, bound by tc_gan/clib.py) plus batched entry
 * points for the same hot path.  Plain pointers and sizes only; no torch types.
 * Every function is re-entrant and may be called concurrently from several
 * host threads (the reference calls its solver from a thread pool with the GIL
 * released, tc_gan/ssnode.py:455-460).
 *
 * Citations are into /root/reference/tc_gan.
 */
#ifndef SSNODE_B200_H
#define SSNODE_B200_H

#if defined(__GNUC__)
#define SSN_API __attribute__((visibility("default")))
#else
#define SSN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * 1. Reference ABI, unchanged (clib.py:16-33  <->  ext/ssnode.c:10-53,55-62).
 *
 * One network x one stimulus, host pointers, float64.  `r0` is the initial
 * state and receives the result; `r1` is 2N doubles of scratch (on return
 * code 0 it also holds the fixed point, ext/ssnode.c:176-178).
 * Return: 0 converged, 1 max_iter reached, 2 rate_hard_bound reached
 * (power/linear only), 1000 + cudaError_t when the GPU call itself failed
 * (the reference reserves >900 for library errors, ssnode.py:267-268).
 * The solve runs on the GPU in float64 (kernel ssn_fp64_kernel).
 * ---------------------------------------------------------------------- */
SSN_API int solve_dynamics_asym_power_euler(int N, double *W, double *ext, double k, double n,
                                    double *r0, double *r1, double tau_E, double tau_I,
                                    double dt, int max_iter, double atol,
                                    double rate_soft_bound, double rate_hard_bound);
SSN_API int solve_dynamics_asym_linear_euler(int N, double *W, double *ext, double k, double n,
                                     double *r0, double *r1, double tau_E, double tau_I,
                                     double dt, int max_iter, double atol,
                                     double rate_soft_bound, double rate_hard_bound);
SSN_API int solve_dynamics_asym_tanh_euler(int N, double *W, double *ext, double k, double n,
                                   double *r0, double *r1, double tau_E, double tau_I,
                                   double dt, int max_iter, double atol,
                                   double rate_soft_bound, double rate_hard_bound);

/* Scalar helpers (ext/ssnode.c:10-53); host functions sharing their source
 * with the device code.  `dot` is exported by the reference but never bound. */
SSN_API double dot(int dim, const double *x, const double *y);
SSN_API double rate_to_volt(double rate, double k, double n);
SSN_API double io_pow(double v, double r0, double r1, double v0, double k, double n);
SSN_API double io_alin(double v, double r0, double r1, double v0, double k, double n);
SSN_API double io_atanh(double v, double r0, double r1, double v0, double k, double n);

/* ------------------------------------------------------------------------
 * 2. Batched entry points (new).  They replace the Python thread pool of
 * ssnode.find_fixed_points_parallel (ssnode.py:423-510), the numpy weight
 * generation (weight_gen.py:13-26), the Theano implicit gradient
 * (gradient_expressions/SS_grad.py:17-76 + make_w_batch.py:36-121 +
 * run/gan.py:902-911) and the Theano Euler unroll with its autodiff
 * (networks/ssn.py:555-576, 598-633).
 * ---------------------------------------------------------------------- */

enum { SSN_IO_POWER = 0, SSN_IO_LINEAR = 1, SSN_IO_TANH = 2 };   /* io_type */
enum { SSN_MEM_HOST = 0, SSN_MEM_DEVICE = 1 };                    /* where the arrays live */
enum { SSN_W_DENSE = 0, SSN_W_FROM_Z = 1 };                       /* `w` holds W, or z (W built on chip) */

/* Solver knobs of ssnode.fixed_point (ssnode.py:159-165). */
typedef struct ssn_solver {
    int    io_type;
    int    max_iter;
    double k, n;
    double tau_E, tau_I, dt;
    double atol;
    double rate_soft_bound;
    double rate_hard_bound;     /* power/linear: the caller passes rate_stop_at here (ssnode.py:241-242) */
} ssn_solver;

/* Generator parameters, row-major 2x2 each: [EE, EI, IE, II]. */
typedef struct ssn_jds { double J[4], D[4], S[4]; } ssn_jds;

/*
 * Fixed points of nz networks x nb stimuli, each solve started from r_init
 * (NULL = zeros).  FP32 FFMA contraction with W resident in (cluster-
 * distributed) shared memory, float64 state update; `precise != 0` selects
 * the all-float64 kernel instead.
 *
 *   w        float32 [nz][2N][2N]   W (w_kind = SSN_W_DENSE) or z (SSN_W_FROM_Z)
 *   jds      used only with SSN_W_FROM_Z
 *   ext      float32 [nb][2N]  (ext_per_network = 0)  or [nz][nb][2N] (= 1)
 *   r_init   float32 [nz][nb][2N] or NULL
 *   R        float32 [nz][nb][2N]   out: final state of every solve
 *   status   int32   [nz][nb]       out: 0 / 1 / 2 as the reference's return code;
 *                                   a converged but non-finite state is reported
 *                                   as 1 (ssnode.py:257-262)
 *   iters    int32   [nz][nb]       out: Euler sweeps performed (may be NULL)
 *   mem      SSN_MEM_HOST: every array is host memory (copied in and out here);
 *            SSN_MEM_DEVICE: device pointers, work is enqueued on `stream`
 *            (a cudaStream_t, NULL = default stream) and NOT synchronised.
 * Returns 0, or 1000 + cudaError_t, or -1 for an unsupported shape.
 */
SSN_API int ssn_fixed_point_batch(const ssn_solver *solver, int nz, int nb, int n_sites,
                          int w_kind, const float *w, const ssn_jds *jds,
                          const float *ext, int ext_per_network, const float *r_init,
                          float *R, int *status, int *iters,
                          int precise, int mem, void *stream);

/* Same, float64 host arrays exactly as the reference passes them
 * (W [nz][2N][2N], ext [nb][2N], R [nz][nb][2N]); used by ssnode.find_fixed_points.
 * precise = 0: W and ext are rounded to float32 on the device for the fast kernel. */
SSN_API int ssn_fixed_point_batch_f64(const ssn_solver *solver, int nz, int nb, int n_sites,
                              const double *W, const double *ext, const double *r_init,
                              double *R, int *status, int *iters, int precise);

/*
 * The same solve for a LIST of float64 host matrices, one pointer per network (the reference's callers hold
 * one numpy array per network, ssnode.py:436-447; nothing is concatenated on the caller's side).
 *   items    nz pointers to row-major [2N][2N] matrices, float64 (items_f32 = 0) or float32 (= 1): W
 *            (w_kind = SSN_W_DENSE), or z with W built on chip from jds (SSN_W_FROM_Z: a caller that owns
 *            J, D, S ships z and skips its own W construction)
 *   ext      float64 [nb][2N];  r_init float64 [nz][nb][2N] or NULL;  R float64 [nz][nb][2N] out
 * The matrices are rounded to float32 by `host_threads` host threads (<= 0: up to 16) into pinned staging,
 * slab by slab, while the GPU solves the previous slab (three pinned slabs, two streams): the fast FP32 kernel
 * with float64 state, i.e. ssn_fixed_point_batch_f64(precise = 0) without its host-side bottlenecks.
 */
SSN_API int ssn_fixed_point_batch_ptrs(const ssn_solver *solver, int nz, int nb, int n_sites, int w_kind,
                               const void *const *items, int items_f32, const ssn_jds *jds,
                               const double *ext, const double *r_init, double *R, int *status, int *iters,
                               int host_threads);
/* dst[i * bytes ...] = the `bytes` bytes at items[i] (i < n), copied by several host threads: stacks the kept
 * per-network arrays into the Zs array find_fixed_points returns (ssnode.py:503). */
SSN_API int ssn_host_gather(const void *const *items, int n, size_t bytes, void *dst, int host_threads);

/*
 * Implicit-function-theorem generator gradient at the fixed points:
 *   (I - W^T Phi) mu = g,   dL/dW = (Phi mu) r^T,   dL/dtheta = <dL/dW, dW/dtheta>
 * with Phi = diag f'(W r + ext) (SS_grad.py:45-59) and g = dL/dr.  The adjoint
 * system is solved by restarted GMRES(16), the eight stimuli of a panel in lockstep,
 * restarted from the true residual (environment SSN_IFT=damped selects the damped
 * iteration mu <- mu + eps (g - mu + W^T Phi mu), eps = dt/tau, of the first version;
 * its stopping rule is max|d mu| < rtol max|g|).
 *
 *   z        float32 [nz][2N][2N]  (W is rebuilt on chip from z and jds)
 *   R, g     float32 [nz][nb][2N]  fixed points and dL/dr
 *   grad     float64 [12] out: dL/dJ[4], dL/dD[4], dL/dS[4] summed over z and b
 *            (device pointer when mem = SSN_MEM_DEVICE; accumulated with atomics,
 *            zeroed by this call)
 *   mu       float32 [nz][nb][2N] out, may be NULL
 *   status   int32 [nz][nb] out: 0 converged (or stalled at the FP32 floor of the residual within 16 rtol);
 *            1 tolerance not reached (max_iter sweeps, or the true residual stopped shrinking above 16 rtol:
 *            ill-conditioned system -- mu is still the best iterate);
 *            iters = contractions with W^T spent on the solve; both may be NULL
 *   rtol     stop when |g - (I - W^T Phi) mu|_2 <= rtol |g|_2 per (network, stimulus); <= 0 selects 1e-6
 *   grad_ext float32 [nz][nb][2N] out, may be NULL: dL/d ext = Phi mu (the gradient w.r.t. the stimulus
 *            input, needed by the heterogeneous-input generators, networks/ssn.py:645-727)
 */
SSN_API int ssn_ift_gradient_batch(const ssn_solver *solver, int nz, int nb, int n_sites,
                           const float *z, const ssn_jds *jds, const float *ext,
                           int ext_per_network, const float *R, const float *g,
                           double rtol, double *grad, float *mu, int *status, int *iters,
                           float *grad_ext, int mem, void *stream);

/*
 * Unrolled Euler dynamics r_{t+1} = (1-eps) r_t + eps f(W r_t + I), r_0 = 0,
 * t = 0..seqlen-1 (networks/ssn.py:555-576), with the outputs of
 * EulerSSNModel (networks/ssn.py:619-633) accumulated over the kept steps
 * t >= skip_steps:
 *   time_avg   float32 [nz][nb][2N]
 *   penalties  float64 [2]: sum (r_{t+1}-r_t)^2 and sum relu(r_t - threshold)
 *              (un-normalised sums; the caller divides by the element counts)
 *   traj       float32 [nz][seqlen][nb][pitch] or NULL: states r_1..r_seqlen
 *   gain       float32 [nz][seqlen][nb][pitch] or NULL: eps*f'(W r_t + I), t = 0..seqlen-1
 *              (pitch = ssn_traj_pitch(n_sites) = 2N rounded up to a multiple of 4 floats: 16-byte rows, which
 *              the TMA loads of the backward pass need; the pad columns are never read)
 * eps_E = dt/tau_E, eps_I = dt/tau_I from `solver`.  Device pointers only.
 */
SSN_API int ssn_euler_forward(const ssn_solver *solver, int nz, int nb, int n_sites,
                      const float *z, const ssn_jds *jds, const float *ext,
                      int ext_per_network, int seqlen, int skip_steps,
                      double rate_penalty_threshold,
                      float *time_avg, double *penalties, float *traj, float *gain,
                      void *stream);

/*
 * BPTT through ssn_euler_forward: given dL/d time_avg [nz][nb][2N] and the
 * scalar weights dL/d(dynamics sum), dL/d(rate sum), returns dL/dJ, dL/dD,
 * dL/dS in grad[12] (float64, device, zeroed here).  `traj` and `gain` are the
 * arrays the forward call stored; `adj` is scratch of the same size as traj (all three with rows of
 * ssn_traj_pitch(n_sites) floats).  The parameter-gradient contraction sum_k adj_k traj_k^T runs on the
 * tensor cores (tcgen05 kind::tf32 with a 3-term split, TMA operand loads).
 * grad_ext (float32 [nz][nb][2N], may be NULL) receives dL/d ext = sum_t gain[t] * lambda_{t+1}.
 * w_dev (device float32 [2], may be NULL): when given, the kernel multiplies w_dyn and w_rate by w_dev[0] and
 * w_dev[1] read on the device, so that upstream scalar gradients need no device-to-host copy.
 */
SSN_API int ssn_euler_backward(const ssn_solver *solver, int nz, int nb, int n_sites,
                       const float *z, const ssn_jds *jds,
                       int seqlen, int skip_steps, double rate_penalty_threshold,
                       const float *grad_time_avg, double w_dyn, double w_rate, const float *w_dev,
                       const float *traj, const float *gain, float *adj,
                       double *grad, float *grad_ext, void *stream);

/* The parameter-gradient contraction of ssn_euler_backward alone: grad[12] (device float64, zeroed here) =
 * < sum_k adj_k traj_k^T, dW/d(J, D, S) > over k = (t, b); adj, traj [nz][seqlen][nb][pitch] device float32. */
SSN_API int ssn_bptt_param_grad(int nz, int nb, int n_sites, int seqlen, const float *adj, const float *traj,
                        const float *z, const ssn_jds *jds, double *grad, void *stream);

/*
 * Probes: tuning_curve[i][b] = rates[model_ids[i]][b][probes[i]] for i < batch -- the gather of
 * networks/cwgan.py:96-99 (ConditionalProber; FixedProber networks/ssn.py:838-851 is the special case
 * model_ids = repeat(arange(nz)), probes = tile(fixed probes)) -- and its gradient, the scatter-add of
 * grad_out [batch][nb] into grad_rates [nz][nb][2N] (zeroed here), which is the dL/dr array
 * ssn_euler_backward / ssn_ift_gradient_batch consume.  Device pointers; model_ids and probes are int32.
 */
SSN_API int ssn_probe_gather(const float *rates, const int *model_ids, const int *probes,
                     int batch, int nz, int nb, int n_sites, float *out, void *stream);
SSN_API int ssn_probe_scatter(const float *grad_out, const int *model_ids, const int *probes,
                      int batch, int nz, int nb, int n_sites, float *grad_rates, void *stream);

/* Build W [nz][2N][2N] (float32) from z on the device (weight_gen.py:13-26). */
SSN_API int ssn_generate_weight(int nz, int n_sites, const float *z, const ssn_jds *jds,
                        float *W, int mem, void *stream);

/* Introspection. */
SSN_API int ssn_device_count(void);                 /* 0 when no usable GPU */
SSN_API const char *ssn_last_error(void);           /* thread-local text of the last failure */
SSN_API int ssn_kernel_launches(void);              /* kernels launched by this library so far (process-wide) */
/* clusters the fixed-point kernel keeps resident for a given size, and its cluster width */
SSN_API int ssn_fixed_point_occupancy(int n_sites, int *cluster_size, int *resident_clusters);

/* floats per row of the BPTT scratch arrays (traj, gain, adj): 2N rounded up to a multiple of 4 */
SSN_API int ssn_traj_pitch(int n_sites);

/* shape tag of the kernel the FP32 path runs for this size, e.g. "ssn_fp_ws_kernel<NC=14,CW=8,UW=8,TI=7>x8"
 * (kernel<template shape> x cluster width): keys the committed ncu figures in profiles/ */
SSN_API int ssn_fixed_point_kernel_name(int n_sites, char *buf, int cap);

/* Per-kernel device time: after ssn_profile_enable(1) every launch of this library is bracketed by CUDA
 * events on its own stream; ssn_profile_read waits for them and writes "kernel_name total_ms launches" lines
 * (one per kernel, since the previous read) into buf, returning the number of distinct kernels. */
SSN_API int ssn_profile_enable(int on);
SSN_API int ssn_profile_read(char *buf, int cap);

/* measured FP32 FFMA throughput of the current device (dependent-chain probe kernel), TFLOP/s:
 * the roofline denominator of the fixed-point kernel */
SSN_API int ssn_measure_fp32_peak(double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* SSNODE_B200_H */
