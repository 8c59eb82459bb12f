// kf.cuh -- the 5-state-per-axis LIP Kalman filter of AMR_code_DART/StateFiltering.cpp, one filter per thread.
// Shared by the CUDA kernel (kf_kernels.cu); plain single-precision arithmetic with the association order of the
// reference's Eigen expressions ((A*sigma)*A', (sigma*C')*inv(...), (K*C)*sigma).
#pragma once
#include "../../include/ismpc_b200.h"

namespace ismpc {

struct KfMats { float A[5][5], B[5][2], Cz[3][5], Cxy[3][5]; };

__host__ __device__ inline void kf_build(const ismpc_kf_model_t& m, KfMats& k)   // StateFiltering.cpp:36-61
{
    const float T = m.sampling_time;
    const float A[5][5] = {{1.0f, T, T * T / 2, 0.0f, 0.0f}, {0.0f, 1.0f, T, T, 0.0f}, {0.0f, 0.0f, 1.0f, 0.0f, 0.0f},
                           {0.0f, 0.0f, 0.0f, 1.0f, T}, {0.0f, 0.0f, 0.0f, 0.0f, 1.0f}};
    const float B[5][2] = {{T * T * T / 6, 0.0f}, {T * T / 2, 0.0f}, {T, 0.0f}, {0.0f, T * T / 2}, {0.0f, T}};
    const float Cz[3][5] = {{1.0f, 0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 1.0f, 0.0f, 0.0f}, {0.0f, 0.0f, -m.mass, 1.0f, 0.0f}};
    const float Cxy[3][5] = {{1.0f, 0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 1.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f, 0.0f, 0.0f}};
    for (int i = 0; i < 5; ++i) { for (int j = 0; j < 5; ++j) k.A[i][j] = A[i][j]; for (int j = 0; j < 2; ++j) k.B[i][j] = B[i][j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 5; ++j) { k.Cz[i][j] = Cz[i][j]; k.Cxy[i][j] = Cxy[i][j]; }
}

// state = A state + B [u; 0];  sigma = (A sigma) A' + (B q) B'          (predict_z / predict_xy, StateFiltering.cpp:97-103,115-124)
__host__ __device__ inline void kf_predict(const KfMats& k, const float q[4], float u, float st[5], float sg[25])
{
    float ns[5];
    for (int i = 0; i < 5; ++i) { float a = 0.0f; for (int j = 0; j < 5; ++j) a += k.A[i][j] * st[j]; ns[i] = a + (k.B[i][0] * u + k.B[i][1] * 0.0f); }
    for (int i = 0; i < 5; ++i) st[i] = ns[i];
    float AS[5][5], BQ[5][2];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) { float a = 0.0f; for (int l = 0; l < 5; ++l) a += k.A[i][l] * sg[l * 5 + j]; AS[i][j] = a; }
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 2; ++j) BQ[i][j] = k.B[i][0] * q[0 * 2 + j] + k.B[i][1] * q[1 * 2 + j];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) {
        float a = 0.0f; for (int l = 0; l < 5; ++l) a += AS[i][l] * k.A[j][l];
        sg[i * 5 + j] = a + (BQ[i][0] * k.B[j][0] + BQ[i][1] * k.B[j][1]);
    }
}

__host__ __device__ inline void kf_inv3(const float M[3][3], float R[3][3])
{
    const float c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1], c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2], c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const float det = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02, id = 1.0f / det;
    R[0][0] = c00 * id; R[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id; R[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
    R[1][0] = c01 * id; R[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id; R[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
    R[2][0] = c02 * id; R[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id; R[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
}

// K = (sigma C') inv(R + C sigma C');  state += K (z - (C state + off));  sigma -= (K C) sigma    (update_z / update_xy, :104-112,125-133)
__host__ __device__ inline void kf_update(const float C[3][5], const float R[9], const float z[3], const float off[3], float st[5], float sg[25])
{
    float SC[5][3], S[3][3], Si[3][3], K[5][3];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) { float a = 0.0f; for (int l = 0; l < 5; ++l) a += sg[i * 5 + l] * C[j][l]; SC[i][j] = a; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { float a = 0.0f; for (int l = 0; l < 5; ++l) a += C[i][l] * SC[l][j]; S[i][j] = R[i * 3 + j] + a; }
    kf_inv3(S, Si);
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) K[i][j] = SC[i][0] * Si[0][j] + SC[i][1] * Si[1][j] + SC[i][2] * Si[2][j];
    float inn[3];
    for (int i = 0; i < 3; ++i) { float a = 0.0f; for (int l = 0; l < 5; ++l) a += C[i][l] * st[l]; inn[i] = z[i] - (a + off[i]); }
    for (int i = 0; i < 5; ++i) st[i] = st[i] + (K[i][0] * inn[0] + K[i][1] * inn[1] + K[i][2] * inn[2]);
    float KC[5][5], ns[25];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) KC[i][j] = K[i][0] * C[0][j] + K[i][1] * C[1][j] + K[i][2] * C[2][j];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) { float a = 0.0f; for (int l = 0; l < 5; ++l) a += KC[i][l] * sg[l * 5 + j]; ns[i * 5 + j] = sg[i * 5 + j] - a; }
    for (int e = 0; e < 25; ++e) sg[e] = ns[e];
}

// One FilterWithKalman call (StateFiltering.cpp:77-95).  zmp[2] (nullable): GetZMP() afterwards.
__host__ __device__ inline void kf_step(const ismpc_kf_model_t& m, KfMats& k, ismpc_kf_state_t& s, const ismpc_kf_sample_t& u, float* zmp)
{
    const float offz[3] = {0.0f, 0.0f, -m.g * m.mass}, off0[3] = {0.0f, 0.0f, 0.0f};
    kf_predict(k, m.q_process[2], u.input[2], s.state[2], s.sigma[2]);
    kf_update(k.Cz, m.q_measurement[2], u.meas[2], offz, s.state[2], s.sigma[2]);
    kf_predict(k, m.q_process[0], u.input[0], s.state[0], s.sigma[0]);
    kf_predict(k, m.q_process[1], u.input[1], s.state[1], s.sigma[1]);
    const float f_n = -m.mass * m.g - m.mass * s.state[2][2] + s.state[2][3];      // :127-129
    k.Cxy[2][2] = m.mass * s.state[2][0] / f_n;
    k.Cxy[2][3] = -s.state[2][0] / f_n;
    kf_update(k.Cxy, m.q_measurement[0], u.meas[0], off0, s.state[0], s.sigma[0]);
    kf_update(k.Cxy, m.q_measurement[1], u.meas[1], off0, s.state[1], s.sigma[1]);
    if (zmp) {
        for (int ax = 0; ax < 2; ++ax) { float a = 0.0f; for (int l = 0; l < 5; ++l) a += k.Cxy[2][l] * s.state[ax][l]; zmp[ax] = a; }
    }
}

}  // namespace ismpc
