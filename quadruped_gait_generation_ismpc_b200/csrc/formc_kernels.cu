// formc_kernels.cu -- formulation C kernels: model set-up (MPCSolver constructor work) and the fused tick.
#include "formc.cuh"
#include "launch.h"

namespace ismpc {

// ---------------------------------------------------------------------------------------------------
// Model set-up (MPCSolver::MPCSolver, MPCSolver.cpp:124-160 + the constant part of :252-258).
// H_z = q_p S'S + q_v Sv'Sv + q_u I with S[k][j] = (k-j) dt^2/m, Sv[k][j] = dt/m for j<k (closed forms of
// the reference's matrixPower loops).  Then H_z^-1 via Cholesky, G = S H^-1, M = S H^-1 S'.
// One-time work: simple kernels, no tuning.
// ---------------------------------------------------------------------------------------------------
__global__ void formc_build_H(ismpc_formc_model_t m, double* H)
{
    const int N = m.N;
    const int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || j >= N) return;
    const double c1 = m.dt * m.dt / m.mass, c1v = m.dt / m.mass;
    const int k0 = (i > j ? i : j) + 1;
    double s1 = 0.0;
    for (int k = k0; k < N; ++k) s1 += ((double)(k - i) * c1) * ((double)(k - j) * c1);
    double s2 = (double)(N - k0 > 0 ? N - k0 : 0) * c1v * c1v;
    H[(size_t)i * N + j] = m.q_p * s1 + m.q_v * s2 + (i == j ? m.q_u : 0.0);
}

// In-place lower Cholesky of an N x N row-major matrix in global memory by ONE CTA (right-looking).
__global__ void formc_cholesky(int N, double* A, int* info)
{
    __shared__ double piv;
    for (int k = 0; k < N; ++k) {
        if (threadIdx.x == 0) {
            double d = A[(size_t)k * N + k];
            if (!(d > 0.0)) { *info = k + 1; d = 1.0; }
            piv = sqrt(d);
            A[(size_t)k * N + k] = piv;
        }
        __syncthreads();
        const double p = piv;
        for (int i = k + 1 + threadIdx.x; i < N; i += blockDim.x) A[(size_t)i * N + k] /= p;
        __syncthreads();
        // trailing update: A[i][j] -= L[i][k] L[j][k] for k < j <= i
        const int rem = N - k - 1;
        for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
            int i = k + 1 + e / rem, j = k + 1 + e % rem;
            if (j <= i) A[(size_t)i * N + j] -= A[(size_t)i * N + k] * A[(size_t)j * N + k];
        }
        __syncthreads();
    }
}

// Linv: thread per column c solves L x = e_c (x lower part only).  Linv stored row-major N x N (upper part 0).
__global__ void formc_tri_inverse(int N, const double* L, double* Linv)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    for (int i = 0; i < N; ++i) {
        double s = (i == c) ? 1.0 : 0.0;
        if (i < c) { Linv[(size_t)i * N + c] = 0.0; continue; }
        for (int k = c; k < i; ++k) s -= L[(size_t)i * N + k] * Linv[(size_t)k * N + c];
        Linv[(size_t)i * N + c] = s / L[(size_t)i * N + i];
    }
}

// Hinv = Linv' Linv
__global__ void formc_hinv(int N, const double* Linv, double* Hinv)
{
    const int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || j >= N) return;
    const int k0 = i > j ? i : j;
    double s = 0.0;
    for (int k = k0; k < N; ++k) s += Linv[(size_t)k * N + i] * Linv[(size_t)k * N + j];
    Hinv[(size_t)i * N + j] = s;
}

// G[k][i] = sum_{j<k} (k-j) c1 Hinv[j][i]
__global__ void formc_G(ismpc_formc_model_t m, const double* Hinv, double* G)
{
    const int N = m.N;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N || i >= N) return;
    const double c1 = m.dt * m.dt / m.mass;
    double s = 0.0;
    for (int j = 0; j < k; ++j) s += ((double)(k - j) * c1) * Hinv[(size_t)j * N + i];
    G[(size_t)k * N + i] = s;
}
// M[k][l] = sum_{j<l} (l-j) c1 G[k][j]
__global__ void formc_M(ismpc_formc_model_t m, const double* G, double* M)
{
    const int N = m.N;
    const int k = blockIdx.y * blockDim.y + threadIdx.y, l = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N || l >= N) return;
    const double c1 = m.dt * m.dt / m.mass;
    double s = 0.0;
    for (int j = 0; j < l; ++j) s += ((double)(l - j) * c1) * G[(size_t)k * N + j];
    M[(size_t)k * N + l] = s;
}

// Projector tables of a prepared gait: one CTA per mpcIter m builds
//   P_m = H^-1 - H^-1[:,K] (H^-1_KK)^-1 H^-1[K,:]      (K = flight-phase columns at m; rows/columns K set to exactly 0)
// Dynamic shared memory: Sk[ne*2ne] (Gauss-Jordan on [H^-1_KK | I]) + W[ne*N] = (H^-1_KK)^-1 H^-1[K,:].
__global__ void formc_build_P(int N, int S, int F, const double* __restrict__ Hinv, double* __restrict__ P, int* info)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = blockIdx.x;
    int ne, c_lo;
    formc_flight_range(N, S, F, m, c_lo, ne);
    double* Pm = P + (size_t)m * N * N;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (ne == 0) {
        for (size_t e = tid; e < (size_t)N * N; e += nt) Pm[e] = Hinv[e];
        return;
    }
    double* Sk = reinterpret_cast<double*>(smem_raw);            // ne x 2ne
    double* W = Sk + (size_t)ne * 2 * ne;                        // ne x N
    const int LD = 2 * ne;
    for (int e = tid; e < ne * ne; e += nt) {
        const int r = e / ne, c = e - r * ne;
        Sk[r * LD + c] = Hinv[(size_t)(c_lo + r) * N + c_lo + c];
        Sk[r * LD + ne + c] = r == c ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int p = 0; p < ne; ++p) {                               // SPD block: no pivoting needed
        const double piv = Sk[p * LD + p];
        if (!(piv > 0.0) && tid == 0) *info = 1000 + m;
        __syncthreads();
        for (int c = tid; c < LD; c += nt) if (c != p) Sk[p * LD + c] /= piv;
        __syncthreads();
        if (tid == 0) Sk[p * LD + p] = 1.0;
        for (int e = tid; e < ne * LD; e += nt) {
            const int r = e / LD, c = e - r * LD;
            if (r != p && c != p) Sk[r * LD + c] -= Sk[r * LD + p] * Sk[p * LD + c];
        }
        __syncthreads();
        for (int r = tid; r < ne; r += nt) if (r != p) Sk[r * LD + p] = 0.0;
        __syncthreads();
    }
    for (int e = tid; e < ne * N; e += nt) {                     // W = Sk^-1 Hinv[K,:]
        const int r = e / N, j = e - r * N;
        double acc = 0.0;
        for (int k = 0; k < ne; ++k) acc += Sk[r * LD + ne + k] * Hinv[(size_t)(c_lo + k) * N + j];
        W[e] = acc;
    }
    __syncthreads();
    for (size_t e = tid; e < (size_t)N * N; e += nt) {
        const int i = (int)(e / N), j = (int)(e - (size_t)i * N);
        double v = 0.0;
        if (!(i >= c_lo && i < c_lo + ne) && !(j >= c_lo && j < c_lo + ne)) {
            v = Hinv[e];
            for (int k = 0; k < ne; ++k) v -= Hinv[(size_t)i * N + c_lo + k] * W[(size_t)k * N + j];
        }
        Pm[e] = v;
    }
}

// Returns 0, a cudaError, or -1 if the flight-phase block does not fit in shared memory (then no tables are built).
int formc_prepare_gait_launch(int N, int S, int F, const double* Hinv, double* P, int* d_info, cudaStream_t st,
                              long long* launches)
{
    int ne_max = F > S + F ? F : S + F;
    if (ne_max > N) ne_max = N;
    const size_t smem = ((size_t)ne_max * 2 * ne_max + (size_t)ne_max * N) * sizeof(double);
    if (smem > (size_t)227 * 1024) return -1;
    cudaError_t e = cudaFuncSetAttribute(formc_build_P, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    formc_build_P<<<S + F, 256, smem, st>>>(N, S, F, Hinv, P, d_info);
    *launches += 1;
    return (int)cudaGetLastError();
}

int formc_setup_launch(const ismpc_formc_model_t& m, double* work /*3 N^2*/, double* Hinv, double* G, double* M,
                       int* d_info, cudaStream_t st, long long* launches)
{
    const int N = m.N;
    double* H = work; double* Linv = work + (size_t)N * N;
    dim3 b(16, 16), gr((N + 15) / 16, (N + 15) / 16);
    formc_build_H<<<gr, b, 0, st>>>(m, H);
    formc_cholesky<<<1, 1024, 0, st>>>(N, H, d_info);
    formc_tri_inverse<<<(N + 63) / 64, 64, 0, st>>>(N, H, Linv);
    formc_hinv<<<gr, b, 0, st>>>(N, Linv, Hinv);
    formc_G<<<gr, b, 0, st>>>(m, Hinv, G);
    formc_M<<<gr, b, 0, st>>>(m, G, M);
    *launches += 6;
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Fused tick: one CTA per instance.
// ---------------------------------------------------------------------------------------------------
#ifndef FORMC_MIN_CTAS
#define FORMC_MIN_CTAS 7
#endif
__global__ void __launch_bounds__(FORMC_THREADS, FORMC_MIN_CTAS)
formc_tick_kernel(FormCArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FormCShared sm;
    formc_carve(smem_raw, a.model.N, sm);
    if (threadIdx.x == 0) { mbar_init(sm.bar, 1); mbar_fence_init(); }
    __syncthreads();
    uint32_t parity = 0;
    const int N = a.model.N;
    for (int inst = blockIdx.x; inst < a.n; inst += gridDim.x) {
        const ismpc_state_t st = a.state[inst];
        const ismpc_walk_t wk = a.walk[inst];
        const ismpc_formc_inst_t in = formc_checked_inst(a.inst[inst], a.plan_rows);
        formc_tick<false>(sm, a.model, a.T, st, wk, in, a.plan, a.out + inst,
                   a.primal ? a.primal + (size_t)inst * 3 * N : nullptr,
                   a.active ? a.active + (size_t)inst * 3 * N : nullptr, parity);
        __syncthreads();
    }
}

// Cluster-per-QP variant for long horizons: a cluster of CS CTAs owns one instance (see formc_tick<true>).
__global__ void __launch_bounds__(FORMC_THREADS)
formc_tick_cluster_kernel(FormCArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto cl = cooperative_groups::this_cluster();
    const unsigned cs = cl.num_blocks(), cr = cl.block_rank();
    FormCShared sm;
    formc_carve(smem_raw, a.model.N, sm);
    if (threadIdx.x == 0) { mbar_init(sm.bar, 1); mbar_fence_init(); }
    __syncthreads();
    cl.sync();                                   // peers' shared memory is live before anyone stores into it
    uint32_t parity = 0;
    const int N = a.model.N;
    const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
    for (int inst = cluster_id; inst < a.n; inst += n_clusters) {
        const ismpc_state_t st = a.state[inst];
        const ismpc_walk_t wk = a.walk[inst];
        const ismpc_formc_inst_t in = formc_checked_inst(a.inst[inst], a.plan_rows);
        const bool lead = cr == 0;
        formc_tick<true>(sm, a.model, a.T, st, wk, in, a.plan, lead ? a.out + inst : nullptr,
                         lead && a.primal ? a.primal + (size_t)inst * 3 * N : nullptr,
                         lead && a.active ? a.active + (size_t)inst * 3 * N : nullptr, parity);
        cl.sync();                               // nobody stores the next instance's slice while a peer still reads
    }
}

struct FormCRolloutArgs {
    FormCArgs base;
    ismpc_state_t* state_io;
    ismpc_walk_t* walk_io;
    const ismpc_push_t* push;   // nullable
    int n_ticks;
    double* traj;               // nullable, n x n_ticks x 6
    int32_t* status;            // nullable
    int32_t* trace;             // nullable, n x n_ticks: status of every tick
};

// Closed loop: the CTA keeps its instance and advances it n_ticks times (Controller::update bookkeeping,
// Controller.cpp:297-302 with the footstep switch enabled, :503-504).
__global__ void __launch_bounds__(FORMC_THREADS)
formc_rollout_kernel(FormCRolloutArgs ra)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ ismpc_formc_out_t s_out;
    FormCShared sm;
    const FormCArgs& a = ra.base;
    formc_carve(smem_raw, a.model.N, sm);
    if (threadIdx.x == 0) { mbar_init(sm.bar, 1); mbar_fence_init(); }
    __syncthreads();
    uint32_t parity = 0;
    for (int inst = blockIdx.x; inst < a.n; inst += gridDim.x) {
        ismpc_state_t st = ra.state_io[inst];
        ismpc_walk_t wk = ra.walk_io[inst];
        const ismpc_formc_inst_t in = formc_checked_inst(a.inst[inst], a.plan_rows);
        ismpc_push_t pu; pu.fs = 0; pu.ct0 = 0; pu.ct1 = 0; pu.ax = 0.0; pu.ay = 0.0; pu.reserved = 0;
        if (ra.push) pu = ra.push[inst];
        int acc_status = 0;
        const double* plan_t = a.plan + (size_t)in.plan_first_row * 4;
        for (int tick = 0; tick < ra.n_ticks; ++tick) {
            // footstep switch (Controller.cpp:297-302, enabled)
            if (wk.footstep_counter >= 0 && wk.footstep_counter < in.n_steps &&
                wk.sim_time >= plan_t[(size_t)wk.footstep_counter * 4 + 3] - 1.0) {
                wk.control_iter = 0; wk.mpc_iter = 0; wk.footstep_counter += 1; wk.support_foot = !wk.support_foot;
            }
            if (tick >= pu.ct0 && tick < pu.ct1) {     // impulsive push (quad_as_bip_bang.m:104-114)
                st.com_vel[0] += a.model.dt * pu.ax; st.com_vel[1] += a.model.dt * pu.ay;
            }
            formc_tick<false>(sm, a.model, a.T, st, wk, in, a.plan, &s_out, nullptr, nullptr, parity);
            __syncthreads();
            st = s_out.next;
            acc_status |= s_out.status;
            if (ra.trace && threadIdx.x == 0) ra.trace[(size_t)inst * ra.n_ticks + tick] = s_out.status;
            if (ra.traj && threadIdx.x < 6) {
                double v = threadIdx.x < 3 ? st.com_pos[threadIdx.x] : st.com_vel[threadIdx.x - 3];
                ra.traj[((size_t)inst * ra.n_ticks + tick) * 6 + threadIdx.x] = v;
            }
            wk.control_iter += 1;                                              // Controller.cpp:503
            wk.mpc_iter = (int)floor(wk.control_iter * a.model.dtc / a.model.dt); // Controller.cpp:504
            wk.sim_time += 1.0;                                                // Controller.cpp:310 (sim frames)
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            ra.state_io[inst] = st; ra.walk_io[inst] = wk;
            if (ra.status) ra.status[inst] = acc_status;
        }
        __syncthreads();
    }
}

// CTAs of the cluster kernel that one SM keeps resident at horizon N (for the cluster-size policy in api.cu).
int formc_cluster_ctas_per_sm(int N)
{
    size_t smem = formc_smem_bytes(N);
    cudaFuncSetAttribute(formc_tick_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, formc_tick_cluster_kernel, FORMC_THREADS, smem) != cudaSuccess) nb = 1;
    return nb > 0 ? nb : 1;
}

int formc_tick_launch(const FormCArgs& a, int grid, int cluster_size, cudaStream_t st)
{
    size_t smem = formc_smem_bytes(a.model.N);
    if (smem > 48 * 1024) {     // per-device function attribute, needed only beyond the default limit (N > ~330); no process-wide cache
        cudaFuncSetAttribute(formc_tick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(formc_tick_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (cluster_size <= 1) {
        formc_tick_kernel<<<grid, FORMC_THREADS, smem, st>>>(a);
        return (int)cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid * cluster_size); cfg.blockDim = dim3(FORMC_THREADS);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, formc_tick_cluster_kernel, a);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

int formc_rollout_launch(const FormCArgs& a, ismpc_state_t* state_io, ismpc_walk_t* walk_io, const ismpc_push_t* push,
                         int n_ticks, double* traj, int32_t* status, int32_t* trace, int grid, cudaStream_t st)
{
    size_t smem = formc_smem_bytes(a.model.N);
    cudaFuncSetAttribute(formc_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FormCRolloutArgs ra{a, state_io, walk_io, push, n_ticks, traj, status, trace};
    formc_rollout_kernel<<<grid, FORMC_THREADS, smem, st>>>(ra);
    return (int)cudaGetLastError();
}

}  // namespace ismpc
