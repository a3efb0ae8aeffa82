// Probing the simulated rates into tuning curves, and the scatter that is its gradient.
//
// Replaces the Theano advanced indexing of tc_gan/networks/cwgan.py:96-99
// (ConditionalProber: tuning_curve = time_avg[model_ids, :, probes], one probed neuron per batch
// element, varying with the batch element) and tc_gan/networks/ssn.py:838-851 (FixedProber:
// time_avg[:, :, probes] with the same probes for every network, reshaped to
// [nz, nb * n_probes]) -- the output-side boundary of the SSN path.
// The backward of either gather is a scatter-add into dL/d time_avg [nz][nb][2N], which is
// exactly the array the BPTT kernel (ssn_euler_backward) and the implicit-gradient kernel
// (ssn_ift_gradient_batch) take as dL/dr.
#include "ssn_common.cuh"
#include "ssn_launch.h"

namespace ssn {

// out[i][b] = rates[model_ids[i]][b][probes[i]]
__global__ void ssn_probe_gather_kernel(const float *__restrict__ rates, const int *__restrict__ model_ids,
                                        const int *__restrict__ probes, int batch, int nb, int dim, int nz,
                                        float *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * nb) return;
    const int i = e / nb, b = e - i * nb;
    const int m = model_ids[i], p = probes[i];
    out[e] = (m >= 0 && m < nz && p >= 0 && p < dim) ? __ldg(rates + ((size_t)m * nb + b) * dim + p) : 0.f;
}

// grad_rates[model_ids[i]][b][probes[i]] += grad_out[i][b]   (grad_rates zeroed by the launcher)
__global__ void ssn_probe_scatter_kernel(const float *__restrict__ grad_out, const int *__restrict__ model_ids,
                                         const int *__restrict__ probes, int batch, int nb, int dim, int nz,
                                         float *__restrict__ grad_rates) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * nb) return;
    const int i = e / nb, b = e - i * nb;
    const int m = model_ids[i], p = probes[i];
    if (m >= 0 && m < nz && p >= 0 && p < dim) atomicAdd(grad_rates + ((size_t)m * nb + b) * dim + p, grad_out[e]);
}

int launch_probe_gather(const float *rates, const int *model_ids, const int *probes, int batch, int nz, int nb,
                        int dim, float *out, cudaStream_t stream) {
    const int total = batch * nb;
    if (total <= 0) return 0;
    KernelTimer kt("ssn_probe_gather_kernel", stream);
    ssn_probe_gather_kernel<<<(total + 255) / 256, 256, 0, stream>>>(rates, model_ids, probes, batch, nb, dim, nz, out);
    count_launch();
    return check_cuda(cudaGetLastError(), "ssn_probe_gather");
}

int launch_probe_scatter(const float *grad_out, const int *model_ids, const int *probes, int batch, int nz, int nb,
                         int dim, float *grad_rates, cudaStream_t stream) {
    SSN_CUDA(cudaMemsetAsync(grad_rates, 0, (size_t)nz * nb * dim * sizeof(float), stream));
    const int total = batch * nb;
    if (total <= 0) return 0;
    KernelTimer kt("ssn_probe_scatter_kernel", stream);
    ssn_probe_scatter_kernel<<<(total + 255) / 256, 256, 0, stream>>>(grad_out, model_ids, probes, batch, nb, dim, nz,
                                                                     grad_rates);
    count_launch();
    return check_cuda(cudaGetLastError(), "ssn_probe_scatter");
}

}  // namespace ssn
