// launch.h -- host-side launch entry points of the kernel translation units (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include "../../include/ismpc_b200.h"

namespace ismpc {

struct FormCArgs;
struct FormCTables;
struct FormAArgs;
struct FormCWarpArgs;

int formc_setup_launch(const ismpc_formc_model_t& m, double* work, double* Hinv, double* G, double* M,
                       int* d_info, cudaStream_t st, long long* launches);
int formc_prepare_gait_launch(int N, int S, int F, const double* Hinv, double* P, int* d_info, cudaStream_t st,
                              long long* launches);
int formc_cluster_ctas_per_sm(int N);
int formc_tick_launch(const FormCArgs& a, int grid, int cluster_size, cudaStream_t st);
int formc_rollout_launch(const FormCArgs& a, ismpc_state_t* state_io, ismpc_walk_t* walk_io,
                         const ismpc_push_t* push, int n_ticks, double* traj, int32_t* status, int32_t* trace, int grid,
                         cudaStream_t st);

// warp-per-instance kernels (formc_warp_kernels.cu)
int formc_riccati_launch(const ismpc_formc_model_t& m, int S, int F, int none, double* tab, cudaStream_t st,
                         long long* launches);
int formc_law_launch(const ismpc_formc_model_t& m, int n_pat, const double* ric, double* law, cudaStream_t st, long long* launches);
int formc_warp_supported(int N);
int formc_warp_resident(int N, int sm_count, int res[5]);   // 0 or a cudaError_t
int formc_tick_warp_launch(const FormCWarpArgs& a, int n, const int res[5], int variant, int pdl, int* grid_out, cudaStream_t st);
int formc_rollout_warp_launch(const FormCWarpArgs& a, ismpc_state_t* state_io, ismpc_walk_t* walk_io,
                              const ismpc_push_t* push, int n_ticks, double* traj, int32_t* status, int32_t* trace, int n,
                              const int res[5], int variant, cudaStream_t st);

struct FormALaunchPlan { int R, warps_per_cta, grid, use_pdas, warm_start, kernel; size_t smem, spill_doubles; };
// tuning of the form-A kernels (ismpc_set_option "forma_*"): 0 = default for R and warps_per_cta
struct FormATuning { int R = 0, warps_per_cta = 0, pdas = 1, warm = 1, reg = 1; };
struct FormAOccCache { int per_sm = 0, F3 = -1 /* build of the kernels the entry belongs to */, wpc = 0; size_t smem = 0; };
void forma_plan(const ismpc_forma_model_t& m, int sm_count, long long items, const FormATuning& tune, FormALaunchPlan* p,
                FormAOccCache* occ);
int forma_tick_launch(const FormAArgs& a, const FormALaunchPlan& p, cudaStream_t st);
int forma_rollout_launch(const FormAArgs& a, const FormALaunchPlan& p, ismpc_forma_inst_t* inst_io,
                         double* fs_plan_io, const ismpc_push_t* push, int n_ticks, double* traj, double* pred,
                         int32_t* status, int32_t* trace, cudaStream_t st);
int plan_generate_launch(int n, const ismpc_plan_model_t& m, const ismpc_plan_req_t* req, double* foot_plan, double* center,
                         int rows, cudaStream_t st);
int kf_filter_launch(int n, int n_steps, const ismpc_kf_model_t& m, ismpc_kf_state_t* state, const ismpc_kf_sample_t* samples,
                     float* zmp, cudaStream_t st);
int kf_filter64_launch(int n, int n_steps, const ismpc_kf_model_t& m, ismpc_kf_state64_t* state, const ismpc_kf_sample_t* samples,
                       double* zmp, int joseph, cudaStream_t st);
int feet_place_launch(int n, int n_ticks, const ismpc_feet_model_t& m, const ismpc_feet_inst_t* inst,
                      const int32_t* fs_timing, int timing_len, const double* pred_traj, double* foot_plan,
                      int foot_plan_rows, cudaStream_t st);
int feet_export_launch(int n, const ismpc_feet_model_t& m, const ismpc_feet_inst_t* inst, const double* foot_plan,
                       int foot_plan_rows, int n_steps, int fixed, int swing, double* fl, double* fr, double* rl, double* rr,
                       cudaStream_t st);

int qp_dense_launch(int n, int nV, int nC, const double* H, const double* g, const double* A, const double* lbA,
                    const double* ubA, double* x, double* y, signed char* ws, int32_t* status, int32_t* iters,
                    double* work, int use_dmma, cudaStream_t st);
size_t qp_dense_work_doubles(int n, int nV, int nC);

}  // namespace ismpc
