// Internal launcher prototypes (device pointers, enqueue on `stream`, no sync).
#pragma once
#include "ssn_common.cuh"

namespace ssn {

int launch_fixed_point_f32(const ssn_solver &sv, int nz, int nb, int n_sites, int w_kind, const float *w,
                           const ssn_jds *jds, const float *ext, int ext_per_network, const float *r_init,
                           float *R, int *status, int *iters, int *counter, cudaStream_t stream);
int launch_fixed_point_f64(const ssn_solver &sv, int nz, int nb, int n_sites, const double *W,
                           const double *ext, int ext_per_network, const double *r_init,
                           double *R, int *status, int *iters, bool nonfinite_fixup, int *counter, cudaStream_t stream);
// float64, W resident in the shared memory of a cluster; returns 1 when the shape does not fit
int launch_fixed_point_f64_cluster(const ssn_solver &sv, int nz, int nb, int n_sites, const double *W,
                                   const double *ext, int ext_per_network, const double *r_init,
                                   double *R, int *status, int *iters, int *counter, cudaStream_t stream);
int fixed_point_occupancy(int n_sites, int *cluster_size, int *resident_clusters);
// shape tag of the kernel the FP32 path runs for this size, e.g. "ssn_fp_ws_kernel<NC=14,CW=8,UW=8,TI=7>x8"
int fixed_point_kernel_name(int n_sites, char *buf, int cap);
int ws_kernel_name(const ssn_solver &sv, int n_sites, char *buf, int cap);
// warp-specialised register-resident-W kernel (default); returns 1 when the shape is outside its range
int launch_fixed_point_ws(const ssn_solver &sv, int nz, int nb, int n_sites, int w_kind, const float *w,
                          const ssn_jds *jds, const float *ext, int ext_per_network, const float *r_init,
                          float *R, int *status, int *iters, int *counter, cudaStream_t stream);
int ws_occupancy(const ssn_solver &sv, int n_sites, int *cluster_size, int *resident_clusters);

int launch_ift_gradient(const ssn_solver &sv, int nz, int nb, int n_sites, const float *z, const ssn_jds &jds,
                        const float *ext, int ext_per_network, const float *R, const float *g, double rtol,
                        double *grad, float *mu, int *status, int *iters, float *grad_ext, int *counter,
                        cudaStream_t stream);

int launch_euler_forward(const ssn_solver &sv, int nz, int nb, int n_sites, const float *z, const ssn_jds &jds,
                         const float *ext, int ext_per_network, int seqlen, int skip_steps, double threshold,
                         float *time_avg, double *penalties, float *traj, float *gain, int *counter,
                         cudaStream_t stream);
int launch_euler_backward(const ssn_solver &sv, int nz, int nb, int n_sites, const float *z, const ssn_jds &jds,
                          int seqlen, int skip_steps, double threshold, const float *grad_time_avg,
                          double w_dyn, double w_rate, const float *w_dev, const float *traj, const float *gain, float *adj,
                          double *grad, float *grad_ext, int *counter, cudaStream_t stream);

// tuning_curve[i][b] = rates[model_ids[i]][b][probes[i]] and its scatter-add gradient (zeroes grad_rates first)
int launch_probe_gather(const float *rates, const int *model_ids, const int *probes, int batch, int nz, int nb,
                        int dim, float *out, cudaStream_t stream);
int launch_probe_scatter(const float *grad_out, const int *model_ids, const int *probes, int batch, int nz, int nb,
                         int dim, float *grad_rates, cudaStream_t stream);

// floats per (time step, stimulus) row of the BPTT scratch arrays traj / gain / adj: 2N rounded up to a multiple of 4,
// so that rows are 16-byte aligned (TMA needs 16-byte global strides)
inline int traj_pitch(int n_sites) { return (2 * n_sites + 3) & ~3; }
// dL/dtheta = < sum_k adj_k traj_k^T, dW/dtheta > on the tensor cores (tcgen05 kind::tf32, 3-term split, TMA loads);
// returns 1 when the driver cannot build tensor maps (the caller then uses the FFMA kernel)
int launch_bptt_param_grad_tc(int nz, int n_sites, long long K, int pitch, const float *adj, const float *traj,
                              const float *z, const WeightConst &wc, double *grad, cudaStream_t stream);

// zeroes grad, then the contraction (tensor-core kernel, or the FFMA kernel with SSN_K4B=ffma)
int launch_bptt_param_grad(int nz, int nb, int n_sites, int seqlen, const float *adj, const float *traj, const float *z,
                           const ssn_jds &jds, double *grad, cudaStream_t stream);

int launch_generate_weight(int nz, int n_sites, const float *z, const ssn_jds &jds, float *W, cudaStream_t stream);
int launch_convert_f64_to_f32(const double *src, float *dst, size_t n, cudaStream_t stream);
int launch_convert_f32_to_f64(const float *src, double *dst, size_t n, cudaStream_t stream);

}  // namespace ssn
