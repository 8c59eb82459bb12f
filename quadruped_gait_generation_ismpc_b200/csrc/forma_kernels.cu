// forma_kernels.cu -- formulation A kernels: single tick and closed-loop rollout, one warp per (instance, axis).
#include "forma.cuh"
#include "launch.h"

namespace ismpc {

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v)
{
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void forma_tick_kernel(FormAArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = warp_id(), lane = lane_id();
    const int C = a.model.C, F = a.model.F, n = C + F;
    const size_t wbytes = forma_warp_smem_bytes(C, F, a.L_in_smem);
    const long long item = (long long)blockIdx.x * a.warps_per_cta + warp;
    if (item >= 2LL * a.n) return;
    FormAShared sm;
    forma_carve(smem_raw + wbytes * warp, C, F, a.L_in_smem,
                a.Lwork ? a.Lwork + (size_t)item * forma_L_doubles(C, F) : nullptr, sm);
    const int inst = (int)(item >> 1), axis = (int)(item & 1);
    const ismpc_forma_inst_t in = a.inst[inst];
    const double* plan = a.fs_plan + (size_t)in.plan_first_row * 2;
    const int32_t* ft = a.fs_timing + in.timing_first;
    double s3[3] = {in.st[axis * 3 + 0], in.st[axis * 3 + 1], in.st[axis * 3 + 2]};
    int iters; double kkt;
    int status = forma_tick_axis(sm, a.model, in, s3, in.cur_fs[axis], in.fs_store[axis], in.j, in.fs_counter,
                                 in.cl_first_ramp, plan, ft, axis, &iters, &kkt);
    const double eta = sqrt(a.model.g_eta / in.height);
    const double zd0 = sm.x[0];
    forma_integrate(eta, a.model.dt, s3, zd0);
    ismpc_forma_out_t* o = a.out + inst;
    if (lane < 3) o->st[axis * 3 + lane] = s3[lane];
    for (int f = lane; f < F; f += 32) o->pred_fs[axis * F + f] = sm.x[C + f];
    if (lane == 0) {
        atomicOr(&o->status, status);
        atomicAdd(&o->iters, iters);
        atomic_max_nonneg(&o->kkt_res, kkt);
    }
    if (a.primal) for (int i = lane; i < n; i += 32) a.primal[(size_t)inst * 2 * n + axis * n + i] = sm.x[i];
    if (a.active) {
        signed char* act = a.active + (size_t)inst * 2 * n;
        for (int i = lane; i < C; i += 32) act[axis * C + i] = sm.das.state[i];
        for (int f = lane; f < F; f += 32) act[2 * C + axis * F + f] = sm.das.state[C + f];
    }
}

struct FormARolloutArgs {
    FormAArgs base;
    ismpc_forma_inst_t* inst_io;
    double* plan_io;
    const ismpc_push_t* push;
    int n_ticks;
    double* traj;
    int32_t* status;
};

__global__ void forma_rollout_kernel(FormARolloutArgs ra)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FormAArgs& a = ra.base;
    const int warp = warp_id(), lane = lane_id();
    const int C = a.model.C, F = a.model.F;
    const size_t wbytes = forma_warp_smem_bytes(C, F, a.L_in_smem);
    const long long item = (long long)blockIdx.x * a.warps_per_cta + warp;
    if (item >= 2LL * a.n) return;
    FormAShared sm;
    forma_carve(smem_raw + wbytes * warp, C, F, a.L_in_smem,
                a.Lwork ? a.Lwork + (size_t)item * forma_L_doubles(C, F) : nullptr, sm);
    const int inst = (int)(item >> 1), axis = (int)(item & 1);
    const ismpc_forma_inst_t in = ra.inst_io[inst];
    double* plan = ra.plan_io + (size_t)in.plan_first_row * 2;
    const int32_t* ft = a.fs_timing + in.timing_first;
    const double eta = sqrt(a.model.g_eta / in.height);
    double s3[3] = {in.st[axis * 3 + 0], in.st[axis * 3 + 1], in.st[axis * 3 + 2]};
    double cur = in.cur_fs[axis], store = in.fs_store[axis];
    int j = in.j, fsc = in.fs_counter, first_ramp = in.cl_first_ramp, ct = 0, acc = 0;
    ismpc_push_t pu; pu.fs = -1; pu.ct0 = 0; pu.ct1 = 0; pu.ax = 0.0; pu.ay = 0.0; pu.reserved = 0;
    if (ra.push) pu = ra.push[inst];
    for (int tick = 0; tick < ra.n_ticks; ++tick) {
        if (fsc == pu.fs && ct >= pu.ct0 && ct < pu.ct1) s3[1] += a.model.dt * (axis == 0 ? pu.ax : pu.ay); // bang.m:104-114
        int iters; double kkt;
        acc |= forma_tick_axis(sm, a.model, in, s3, cur, store, j, fsc, first_ramp, plan, ft, axis, &iters, &kkt);
        const double zd0 = sm.x[0], pred = sm.x[C];
        __syncwarp();
        forma_integrate(eta, a.model.dt, s3, zd0);
        if (ra.traj && lane < 3)
            ra.traj[((size_t)inst * ra.n_ticks + tick) * 6 + 2 * lane + axis] = s3[lane];   // x,y,xd,yd,xz,yz
        ct += 1;
        if (fsc + 1 <= in.n_timing && j + 1 >= ft[fsc]) {                                  // bang.m:529
            fsc += 1; cur = pred; store = pred;
            if (fsc >= 2 && fsc <= in.n_fs) {                                              // bang.m:539-556
                const double d = pred - plan[(fsc - 1) * 2 + axis];
                __syncwarp();
                for (int r = lane; r < in.n_fs; r += 32) plan[r * 2 + axis] += d;
                first_ramp = 0;
                __syncwarp();
            }
            ct = 0;
        }
        j += 1;
    }
    ismpc_forma_inst_t* io = ra.inst_io + inst;
    if (lane < 3) io->st[axis * 3 + lane] = s3[lane];
    if (lane == 0) {
        io->cur_fs[axis] = cur; io->fs_store[axis] = store;
        if (axis == 0) { io->j = j; io->fs_counter = fsc; io->cl_first_ramp = first_ramp; }
        if (ra.status) atomicOr(&ra.status[inst], acc);
    }
}

static int forma_configure(FormAArgs& a, size_t* smem_out, int* grid_out, bool rollout)
{
    const int C = a.model.C, F = a.model.F;
    const size_t lim = 227 * 1024;
    size_t per_in = forma_warp_smem_bytes(C, F, true);
    if (per_in <= lim) {
        a.L_in_smem = 1;
        int wpc = (int)(lim / per_in);
        if (wpc > 2) wpc = 2;          // two warps (x and y of one instance) per CTA keeps CTAs small and numerous
        a.warps_per_cta = wpc;
        *smem_out = per_in * wpc;
    } else {
        a.L_in_smem = 0;
        a.warps_per_cta = 2;
        *smem_out = forma_warp_smem_bytes(C, F, false) * 2;
    }
    const long long items = 2LL * a.n;
    *grid_out = (int)((items + a.warps_per_cta - 1) / a.warps_per_cta);
    cudaError_t e = rollout
        ? cudaFuncSetAttribute(forma_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem_out)
        : cudaFuncSetAttribute(forma_tick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem_out);
    return (int)e;
}

size_t forma_Lwork_doubles(const ismpc_forma_model_t& m, int n)
{
    if (forma_warp_smem_bytes(m.C, m.F, true) <= (size_t)227 * 1024) return 0;
    return (size_t)2 * n * forma_L_doubles(m.C, m.F);
}

int forma_tick_launch(const FormAArgs& a_in, cudaStream_t st)
{
    FormAArgs a = a_in;
    size_t smem; int grid;
    int rc = forma_configure(a, &smem, &grid, false);
    if (rc) return rc;
    cudaMemsetAsync(a.out, 0, (size_t)a.n * sizeof(ismpc_forma_out_t), st);
    forma_tick_kernel<<<grid, 32 * a.warps_per_cta, smem, st>>>(a);
    return (int)cudaGetLastError();
}

int forma_rollout_launch(const FormAArgs& a_in, ismpc_forma_inst_t* inst_io, double* fs_plan_io,
                         const ismpc_push_t* push, int n_ticks, double* traj, int32_t* status, cudaStream_t st)
{
    FormAArgs a = a_in;
    size_t smem; int grid;
    int rc = forma_configure(a, &smem, &grid, true);
    if (rc) return rc;
    if (status) cudaMemsetAsync(status, 0, (size_t)a.n * sizeof(int32_t), st);
    FormARolloutArgs ra{a, inst_io, fs_plan_io, push, n_ticks, traj, status};
    forma_rollout_kernel<<<grid, 32 * a.warps_per_cta, smem, st>>>(ra);
    return (int)cudaGetLastError();
}

}  // namespace ismpc
