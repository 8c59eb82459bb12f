// kf_kernels.cu -- batched LIP Kalman filter (AMR_code_DART/StateFiltering.cpp), one filter per thread.
// 3 axes x (5 + 25) floats of state per filter live in registers / local memory for the whole run of n_steps samples;
// samples stream from global memory (n x n_steps records of 48 B).
#include "common.cuh"
#include "kf.cuh"
#include "launch.h"

namespace ismpc {

__global__ void kf_filter_kernel(int n, int n_steps, ismpc_kf_model_t m, ismpc_kf_state_t* state, const ismpc_kf_sample_t* samples, float* zmp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    KfMats k;
    kf_build(m, k);
    ismpc_kf_state_t s = state[i];
    for (int t = 0; t < n_steps; ++t) {
        const ismpc_kf_sample_t u = samples[(size_t)i * n_steps + t];
        float z2[2];
        kf_step(m, k, s, u, z2);
        if (zmp) { zmp[((size_t)i * n_steps + t) * 2] = z2[0]; zmp[((size_t)i * n_steps + t) * 2 + 1] = z2[1]; }
    }
    state[i] = s;
}

// The same filter carried in FP64 (state, covariance, gains; the model and the samples stay the reference's floats), with
// the covariance update in standard or Joseph form.
__global__ void kf_filter64_kernel(int n, int n_steps, ismpc_kf_model_t m, ismpc_kf_state64_t* state, const ismpc_kf_sample_t* samples,
                                   double* zmp, int joseph)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    KfMatsT<double> k;
    kf_build(m, k);
    ismpc_kf_state64_t s = state[i];
    for (int t = 0; t < n_steps; ++t) {
        const ismpc_kf_sample_t u = samples[(size_t)i * n_steps + t];
        double z2[2];
        kf_step_t<double, double>(m, k, s.state, s.sigma, u, z2, joseph);
        if (zmp) { zmp[((size_t)i * n_steps + t) * 2] = z2[0]; zmp[((size_t)i * n_steps + t) * 2 + 1] = z2[1]; }
    }
    state[i] = s;
}

int kf_filter64_launch(int n, int n_steps, const ismpc_kf_model_t& m, ismpc_kf_state64_t* state, const ismpc_kf_sample_t* samples,
                       double* zmp, int joseph, cudaStream_t st)
{
    kf_filter64_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, n_steps, m, state, samples, zmp, joseph);
    return (int)cudaGetLastError();
}

int kf_filter_launch(int n, int n_steps, const ismpc_kf_model_t& m, ismpc_kf_state_t* state, const ismpc_kf_sample_t* samples,
                     float* zmp, cudaStream_t st)
{
    kf_filter_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, n_steps, m, state, samples, zmp);
    return (int)cudaGetLastError();
}

}  // namespace ismpc
