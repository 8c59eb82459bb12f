// kf.cuh -- the 5-state-per-axis LIP Kalman filter of AMR_code_DART/StateFiltering.cpp, one filter per thread.
// Shared by the CUDA kernel (kf_kernels.cu); plain single-precision arithmetic with the association order of the
// reference's Eigen expressions ((A*sigma)*A', (sigma*C')*inv(...), (K*C)*sigma).
#pragma once
#include "../../include/ismpc_b200.h"

namespace ismpc {

template <class T> struct KfMatsT { T A[5][5], B[5][2], Cz[3][5], Cxy[3][5]; };
using KfMats = KfMatsT<float>;

template <class R>
__host__ __device__ inline void kf_build(const ismpc_kf_model_t& m, KfMatsT<R>& k)   // StateFiltering.cpp:36-61
{
    const R T = (R)m.sampling_time;
    const R A[5][5] = {{1, T, T * T / 2, 0, 0}, {0, 1, T, T, 0}, {0, 0, 1, 0, 0},
                       {0, 0, 0, 1, T}, {0, 0, 0, 0, 1}};
    const R B[5][2] = {{T * T * T / 6, 0}, {T * T / 2, 0}, {T, 0}, {0, T * T / 2}, {0, T}};
    const R Cz[3][5] = {{1, 0, 0, 0, 0}, {0, 0, 1, 0, 0}, {0, 0, -(R)m.mass, 1, 0}};
    const R Cxy[3][5] = {{1, 0, 0, 0, 0}, {0, 0, 1, 0, 0}, {1, 0, 0, 0, 0}};
    for (int i = 0; i < 5; ++i) { for (int j = 0; j < 5; ++j) k.A[i][j] = A[i][j]; for (int j = 0; j < 2; ++j) k.B[i][j] = B[i][j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 5; ++j) { k.Cz[i][j] = Cz[i][j]; k.Cxy[i][j] = Cxy[i][j]; }
}

// state = A state + B [u; 0];  sigma = (A sigma) A' + (B q) B'          (predict_z / predict_xy, StateFiltering.cpp:97-103,115-124)
template <class R>
__host__ __device__ inline void kf_predict(const KfMatsT<R>& k, const float q[4], R u, R st[5], R sg[25])
{
    R ns[5];
    for (int i = 0; i < 5; ++i) { R a = 0; for (int j = 0; j < 5; ++j) a += k.A[i][j] * st[j]; ns[i] = a + (k.B[i][0] * u + k.B[i][1] * (R)0); }
    for (int i = 0; i < 5; ++i) st[i] = ns[i];
    R AS[5][5], BQ[5][2];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) { R a = 0; for (int l = 0; l < 5; ++l) a += k.A[i][l] * sg[l * 5 + j]; AS[i][j] = a; }
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 2; ++j) BQ[i][j] = k.B[i][0] * (R)q[0 * 2 + j] + k.B[i][1] * (R)q[1 * 2 + j];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) {
        R a = 0; for (int l = 0; l < 5; ++l) a += AS[i][l] * k.A[j][l];
        sg[i * 5 + j] = a + (BQ[i][0] * k.B[j][0] + BQ[i][1] * k.B[j][1]);
    }
}

template <class T>
__host__ __device__ inline void kf_inv3(const T M[3][3], T R[3][3])
{
    const T c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1], c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2], c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const T det = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02, id = (T)1 / det;
    R[0][0] = c00 * id; R[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id; R[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
    R[1][0] = c01 * id; R[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id; R[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
    R[2][0] = c02 * id; R[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id; R[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
}

// K = (sigma C') inv(R + C sigma C');  state += K (z - (C state + off));  sigma -= (K C) sigma    (update_z / update_xy, :104-112,125-133)
// joseph != 0: the covariance update in Joseph form, sigma = (I - K C) sigma (I - K C)' + K R K' -- algebraically the same
// for the optimal gain, but symmetric and positive semi-definite in floating point by construction (not what the
// reference computes: offered by the FP64 entry point for callers who run the filter for long).
template <class T>
__host__ __device__ inline void kf_update(const T C[3][5], const float R[9], const float z[3], const T off[3], T st[5], T sg[25], int joseph = 0)
{
    T SC[5][3], S[3][3], Si[3][3], K[5][3];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) { T a = 0; for (int l = 0; l < 5; ++l) a += sg[i * 5 + l] * C[j][l]; SC[i][j] = a; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T a = 0; for (int l = 0; l < 5; ++l) a += C[i][l] * SC[l][j]; S[i][j] = (T)R[i * 3 + j] + a; }
    kf_inv3(S, Si);
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) K[i][j] = SC[i][0] * Si[0][j] + SC[i][1] * Si[1][j] + SC[i][2] * Si[2][j];
    T inn[3];
    for (int i = 0; i < 3; ++i) { T a = 0; for (int l = 0; l < 5; ++l) a += C[i][l] * st[l]; inn[i] = (T)z[i] - (a + off[i]); }
    for (int i = 0; i < 5; ++i) st[i] = st[i] + (K[i][0] * inn[0] + K[i][1] * inn[1] + K[i][2] * inn[2]);
    T KC[5][5], ns[25];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) KC[i][j] = K[i][0] * C[0][j] + K[i][1] * C[1][j] + K[i][2] * C[2][j];
    if (!joseph) {
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) { T a = 0; for (int l = 0; l < 5; ++l) a += KC[i][l] * sg[l * 5 + j]; ns[i * 5 + j] = sg[i * 5 + j] - a; }
    } else {
        T IS[5][5], KR[5][3];          // (I - KC) sigma ;  K R
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) { T a = 0; for (int l = 0; l < 5; ++l) a += KC[i][l] * sg[l * 5 + j]; IS[i][j] = sg[i * 5 + j] - a; }
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) KR[i][j] = K[i][0] * (T)R[0 * 3 + j] + K[i][1] * (T)R[1 * 3 + j] + K[i][2] * (T)R[2 * 3 + j];
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) {
            T a = IS[i][j];
            for (int l = 0; l < 5; ++l) a -= IS[i][l] * KC[j][l];
            a += KR[i][0] * K[j][0] + KR[i][1] * K[j][1] + KR[i][2] * K[j][2];
            ns[i * 5 + j] = a;
        }
        for (int i = 0; i < 5; ++i) for (int j = 0; j < i; ++j) { const T m2 = (ns[i * 5 + j] + ns[j * 5 + i]) / 2; ns[i * 5 + j] = m2; ns[j * 5 + i] = m2; }
    }
    for (int e = 0; e < 25; ++e) sg[e] = ns[e];
}

// One FilterWithKalman call (StateFiltering.cpp:77-95) on a state held as T[3][5] / T[3][25].  zmp[2] (nullable): GetZMP().
template <class T, class Z>
__host__ __device__ inline void kf_step_t(const ismpc_kf_model_t& m, KfMatsT<T>& k, T state[3][5], T sigma[3][25], const ismpc_kf_sample_t& u,
                                          Z* zmp, int joseph = 0)
{
    const T offz[3] = {0, 0, -(T)m.g * (T)m.mass}, off0[3] = {0, 0, 0};
    kf_predict(k, m.q_process[2], (T)u.input[2], state[2], sigma[2]);
    kf_update(k.Cz, m.q_measurement[2], u.meas[2], offz, state[2], sigma[2], joseph);
    kf_predict(k, m.q_process[0], (T)u.input[0], state[0], sigma[0]);
    kf_predict(k, m.q_process[1], (T)u.input[1], state[1], sigma[1]);
    const T f_n = -(T)m.mass * (T)m.g - (T)m.mass * state[2][2] + state[2][3];      // :127-129
    k.Cxy[2][2] = (T)m.mass * state[2][0] / f_n;
    k.Cxy[2][3] = -state[2][0] / f_n;
    kf_update(k.Cxy, m.q_measurement[0], u.meas[0], off0, state[0], sigma[0], joseph);
    kf_update(k.Cxy, m.q_measurement[1], u.meas[1], off0, state[1], sigma[1], joseph);
    if (zmp) {
        for (int ax = 0; ax < 2; ++ax) { T a = 0; for (int l = 0; l < 5; ++l) a += k.Cxy[2][l] * state[ax][l]; zmp[ax] = (Z)a; }
    }
}
__host__ __device__ inline void kf_step(const ismpc_kf_model_t& m, KfMats& k, ismpc_kf_state_t& s, const ismpc_kf_sample_t& u, float* zmp)
{
    kf_step_t<float, float>(m, k, s.state, s.sigma, u, zmp, 0);
}

}  // namespace ismpc
