/*
 * oracle_batch.cpp -- batch drivers over the CPU oracle, taking the same POD structs as the product's
 * C ABI (include/ismpc_b200.h is included for its TYPES only).  One cold QP solve per problem per
 * thread, std::thread static partition of instances -- the CPU arm that bench.py times beside the GPU
 *  and the checker the parity tests compare against.
 *
 * TEST INFRASTRUCTURE ONLY -- see ismpc_oracle.h.
 */
#include "ismpc_oracle.h"
#include "../include/ismpc_b200.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#ifdef ORACLE_WITH_QPOASES
extern "C" int oracle_qpoases_solve(int, int, const double*, const double*, const double*, const double*,
                                    const double*, double*, double*, int*, int*);
extern "C" void oracle_qpoases_set_nwsr(int);
#else
extern "C" int oracle_have_qpoases(void) { return 0; }
#endif

static oracle_qp_fn pick_solver(int kind)
{
#ifdef ORACLE_WITH_QPOASES
    if (kind == 1) return oracle_qpoases_solve;
#endif
    if (kind == 0) return oracle_qp_dual_active_set;
    return nullptr;
}

template <class Fn>
static void parallel_for(int n, int nthreads, Fn fn)
{
    if (nthreads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    nthreads = std::min(nthreads, n);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        int lo = (int)((long long)n * t / nthreads), hi = (int)((long long)n * (t + 1) / nthreads);
        th.emplace_back([=]() { for (int i = lo; i < hi; ++i) fn(i); });
    }
    for (auto& t : th) t.join();
}

extern "C" int oracle_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

/* Generic dense batch: n problems of one shape (mirrors ismpc_qp_solve_batch). */
extern "C" int oracle_qp_batch(int solver_kind, int nwsr_cap, int n, int nV, int nC,
                               const double* H, const double* g, const double* A,
                               const double* lbA, const double* ubA,
                               double* x, double* y, int* ws, int* ret, int* nwsr, int nthreads)
{
    oracle_qp_fn solver = pick_solver(solver_kind);
    if (!solver) return -1;
    parallel_for(n, nthreads, [=](int i) {
#ifdef ORACLE_WITH_QPOASES
        oracle_qpoases_set_nwsr(nwsr_cap);
#endif
        int it = 0;
        int rc = solver(nV, nC, H + (size_t)i * nV * nV, g + (size_t)i * nV, A + (size_t)i * nC * nV,
                        lbA + (size_t)i * nC, ubA + (size_t)i * nC, x + (size_t)i * nV,
                        y ? y + (size_t)i * nC : nullptr, ws ? ws + (size_t)i * nC : nullptr, &it);
        if (ret) ret[i] = rc;
        if (nwsr) nwsr[i] = it;
    });
    (void)nwsr_cap;
    return 0;
}

static void formc_params_from_abi(const ismpc_formc_model_t* m, const ismpc_formc_inst_t* in, oracle_formc_params* p)
{
    p->dt = m->dt; p->dtc = m->dtc; p->h = in->com_height; p->mass = m->mass; p->g = m->g;
    p->box_w = in->box_w; p->box_w_init = in->box_w_init;
    p->q_p = m->q_p; p->q_v = m->q_v; p->q_u = m->q_u; p->fz_max = m->fz_max;
    p->N = m->N; p->S = in->S; p->F = in->F_ds;
}

/* Mirrors ismpc_formc_solve_batch.  ret/nwsr: n x 3 (z, x, y).  duals (nullable): n x 3N. */
extern "C" int oracle_formc_batch(int solver_kind, int nwsr_cap, const ismpc_formc_model_t* model, int n,
                                  const ismpc_state_t* state, const ismpc_walk_t* walk,
                                  const ismpc_formc_inst_t* inst, const double* plan_xyzt, int plan_rows,
                                  ismpc_formc_out_t* out, double* primal, int* active, double* duals,
                                  int* ret, int* nwsr, int nthreads)
{
    oracle_qp_fn solver = pick_solver(solver_kind);
    if (!solver) return -1;
    (void)plan_rows;
    int N = model->N;
    parallel_for(n, nthreads, [=](int i) {
#ifdef ORACLE_WITH_QPOASES
        oracle_qpoases_set_nwsr(nwsr_cap);
#endif
        oracle_formc_params p;
        formc_params_from_abi(model, &inst[i], &p);
        oracle_formc_out o;
        double* pr = primal ? primal + (size_t)i * 3 * N : nullptr;
        int* ac = active ? active + (size_t)i * 3 * N : nullptr;
        double* du = duals ? duals + (size_t)i * 3 * N : nullptr;
        int rc = oracle_formc_tick(&p, solver, state[i].com_pos, state[i].com_vel, walk[i].sim_time,
                                   walk[i].mpc_iter, walk[i].control_iter, walk[i].footstep_counter,
                                   plan_xyzt + (size_t)inst[i].plan_first_row * 4, inst[i].n_steps, &o,
                                   pr, pr ? pr + N : nullptr, pr ? pr + 2 * N : nullptr,
                                   ac, ac ? ac + N : nullptr, ac ? ac + 2 * N : nullptr,
                                   du, du ? du + N : nullptr, du ? du + 2 * N : nullptr);
        ismpc_formc_out_t& oo = out[i];
        std::memset(&oo, 0, sizeof(oo));
        oo.next = state[i];
        for (int c = 0; c < 3; ++c) { oo.next.com_pos[c] = o.com_pos[c]; oo.next.com_vel[c] = o.com_vel[c]; }
        oo.zmp_in[0] = o.zmp_in[0]; oo.zmp_in[1] = o.zmp_in[1];
        oo.fz0 = o.fz0; oo.lambda0 = o.lambda0;
        oo.status = (rc == -10) ? ISMPC_ST_WINDOW
                                : ((o.ret[0] ? ISMPC_ST_Z_FAIL : 0) | (o.ret[1] ? ISMPC_ST_X_FAIL : 0) |
                                   (o.ret[2] ? ISMPC_ST_Y_FAIL : 0) |
                                   ((rc != -10 && !(o.lambda0 > 2.0)) ? ISMPC_ST_XY_SKIPPED : 0));
        for (int k = 0; k < 3; ++k) {
            oo.iters[k] = o.nwsr[k];
            if (ret) ret[i * 3 + k] = o.ret[k];
            if (nwsr) nwsr[i * 3 + k] = o.nwsr[k];
        }
    });
    (void)nwsr_cap;
    return 0;
}

static void forma_params_from_abi(const ismpc_forma_model_t* m, const ismpc_forma_inst_t* in, oracle_forma_params* p)
{
    p->dt = m->dt; p->eta = std::sqrt(m->g_eta / in->height); p->wx = in->wx; p->wy = in->wy;
    p->disp_forw = m->disp_forw; p->disp_forw_dummy = m->disp_forw_dummy; p->disp_L = m->disp_L;
    p->Qzdot = m->q_zdot; p->Qfoot = m->q_foot; p->C = m->C; p->P = m->P; p->F = m->F;
}

/* Mirrors ismpc_forma_solve_batch.  active: n x 2(C+F) ints; duals same shape. */
extern "C" int oracle_forma_batch(int solver_kind, int nwsr_cap, const ismpc_forma_model_t* model, int n,
                                  const ismpc_forma_inst_t* inst, const int32_t* fs_timing, int timing_len,
                                  const double* fs_plan, int plan_rows,
                                  ismpc_forma_out_t* out, double* primal, int* active, double* duals,
                                  int* ret, int* nwsr, int nthreads)
{
    oracle_qp_fn solver = pick_solver(solver_kind);
    if (!solver) return -1;
    (void)timing_len; (void)plan_rows;
    int nV = 2 * (model->C + model->F);
    int C = model->C, F = model->F;
    parallel_for(n, nthreads, [=](int i) {
#ifdef ORACLE_WITH_QPOASES
        oracle_qpoases_set_nwsr(nwsr_cap);
#endif
        oracle_forma_params p;
        forma_params_from_abi(model, &inst[i], &p);
        oracle_forma_out o;
        std::vector<int> ft(inst[i].n_timing);
        for (int k = 0; k < inst[i].n_timing; ++k) ft[k] = fs_timing[inst[i].timing_first + k];
        int rc = oracle_forma_tick(&p, solver, inst[i].st, inst[i].cur_fs, inst[i].fs_store, inst[i].j,
                                   inst[i].fs_counter, ft.data(), inst[i].n_timing, inst[i].ds,
                                   fs_plan + (size_t)inst[i].plan_first_row * 2, inst[i].n_fs,
                                   inst[i].cl_first_ramp, &o,
                                   primal ? primal + (size_t)i * nV : nullptr,
                                   active ? active + (size_t)i * nV : nullptr,
                                   duals ? duals + (size_t)i * nV : nullptr);
        ismpc_forma_out_t& oo = out[i];
        std::memset(&oo, 0, sizeof(oo));
        std::memcpy(oo.st, o.st, sizeof(double) * 6);
        if (primal) {
            const double* v = primal + (size_t)i * nV;
            for (int f = 0; f < F; ++f) { oo.pred_fs[f] = v[C + f]; oo.pred_fs[F + f] = v[C + F + C + f]; }
        } else { oo.pred_fs[0] = o.pred_fs[0]; oo.pred_fs[F] = o.pred_fs[1]; }
        oo.status = rc ? ISMPC_ST_QP_FAIL : 0;
        oo.iters = o.nwsr;
        if (ret) ret[i] = o.ret;
        if (nwsr) nwsr[i] = o.nwsr;
    });
    (void)nwsr_cap;
    return 0;
}
