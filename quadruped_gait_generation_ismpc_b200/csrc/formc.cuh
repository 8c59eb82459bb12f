// formc.cuh -- formulation C (what MPCSolver::solve builds): device code for the fused tick.
//
// Reference map (AMR_code_DART/...):
//   MPCSolver.cpp:167-180   midpoint sequence            -> midpoint_value()   (computed on the fly, per window)
//   MPCSolver.cpp:124-160   vertical prediction matrices -> closed forms + the model tables built by
//                                                           formc_setup kernels (H_z^-1, S H^-1, S H^-1 S')
//   MPCSolver.cpp:220-278   stage 1 vertical QP          -> vertical_qp()      (dual active set on tables)
//   MPCSolver.cpp:296-309   stage 2 lambda sequence      -> lambda_and_lip()
//   MPCSolver.cpp:325-389   stage 3 horizontal QP build  -> stability_row() (O(N) suffix-product scan instead
//                                                           of the reference's O(N^2) loop), tails
//   utils.cpp:385-511 / utils.cpp:89-139  solve          -> knapsack_qp()      (exact active-set Newton on the
//                                                           scalar multiplier: H = I, A = [a'; I])
//   MPCSolver.cpp:402-422   integrate                    -> formc_tick() epilogue
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
#include "das.cuh"
#include "../../include/ismpc_b200.h"

namespace ismpc {

constexpr int FORMC_THREADS = 128;
constexpr int FORMC_QMAX = 48;        // max working-set size of the vertical QP (equalities + active rows)
constexpr int FORMC_PLAN_STAGE = 40;  // plan rows staged in shared memory by one bulk copy

struct FormCTables {      // device pointers, N x N row-major each (built once per model)
    const double* Hinv;   // H_z^-1
    const double* G;      // S_bar_z * H_z^-1
    const double* M;      // S_bar_z * H_z^-1 * S_bar_z'
    // Prepared gait (ismpc_formc_prepare_gait): for every mpcIter m of a step of gS + gF ticks, the projector
    //   P_m = H^-1 - H^-1[:,K] (H^-1_KK)^-1 H^-1[K,:],   K = flight-phase columns at mpcIter m (MPCSolver.cpp:223-243),
    // so that f = -P_m F_z is the minimiser under the equalities f_K = 0 in ONE table mat-vec.  Null if not prepared.
    const double* P;      // (gS + gF) x N x N
    int gS, gF;
};

// Flight-phase samples [c_lo, c_lo+ne) of the horizon at mpcIter m (MPCSolver.cpp:223-243, indices as written there).
// The reference fills Aeq_z inside `for (i = 0; i < N; i++)`: with a horizon shorter than S + F only the rows with
// i < N get an entry, the others stay all-zero rows (0 = 0).
__host__ __device__ inline void formc_flight_range(int N, int S, int F, int mpc_iter, int& c_lo, int& ne)
{
    if (mpc_iter < S) {                                      // Aeq_z(i-S, i-mpcIter), i in [S, min(S+F, N))
        const int i_hi = S + F < N ? S + F : N;
        ne = i_hi - S; c_lo = S - mpc_iter;
    } else { ne = S + F - mpc_iter; c_lo = 0; }              // Aeq_z(i,i), i < min(S+F-mpcIter, N)
    if (c_lo < 0) { ne += c_lo; c_lo = 0; }
    if (c_lo + ne > N) ne = N - c_lo;
    if (ne < 0) ne = 0;
}

// A record whose plan rows fall outside the plan table is treated like a window beyond the plan: n_steps = 0 makes the
// tick return ISMPC_ST_WINDOW with the instance untouched (and the rollouts never look a step time up), instead of
// reading out of bounds.
__device__ __forceinline__ ismpc_formc_inst_t formc_checked_inst(ismpc_formc_inst_t in, int plan_rows)
{
    if (in.plan_first_row < 0 || in.n_steps < 0 || (long long)in.plan_first_row + in.n_steps > (long long)plan_rows) {
        in.n_steps = 0; in.plan_first_row = 0;
    }
    return in;
}

struct FormCShared {      // per-CTA shared memory carve-up (all pointers into dynamic smem)
    double *midx, *midy, *midz;            // [2N], [2N], [N]
    double *Fz, *f, *rv, *zdir, *scr;      // [N] each
    double *lam, *chv, *shs, *ssh;         // [N] each: lambda, cosh, sinh/s, s*sinh
    double *avec, *dl;                     // [N] stability row, exp(-dt*eta*i)
    double *plan;                          // [FORMC_PLAN_STAGE*4]
    double *red;                           // [16] small scratch / cross-warp results
    DasWork das;
    uint64_t* bar;
};

__host__ __device__ inline size_t formc_smem_bytes(int N)
{
    size_t d = (size_t)16 * N + FORMC_PLAN_STAGE * 4 + 16 + (FORMC_QMAX * (FORMC_QMAX + 1)) / 2 + 3 * FORMC_QMAX;
    size_t b = d * sizeof(double) + FORMC_QMAX * sizeof(int) + FORMC_QMAX + (size_t)N + 16 /*bar*/;
    return (b + 15) & ~(size_t)15;
}

__device__ inline void formc_carve(unsigned char* base, int N, FormCShared& s)
{
    double* d = reinterpret_cast<double*>(base);
    s.bar = reinterpret_cast<uint64_t*>(d); d += 2;
    s.plan = d; d += FORMC_PLAN_STAGE * 4;
    s.midx = d; d += 2 * N; s.midy = d; d += 2 * N; s.midz = d; d += N;
    s.Fz = d; d += N; s.f = d; d += N; s.rv = d; d += N; s.zdir = d; d += N; s.scr = d; d += N;
    s.lam = d; d += N; s.chv = d; d += N; s.shs = d; d += N; s.ssh = d; d += N;
    s.avec = d; d += N; s.dl = d; d += N;
    s.red = d; d += 16;
    s.das.Js = d; d += (FORMC_QMAX * (FORMC_QMAX + 1)) / 2;
    s.das.Jg = nullptr; s.das.R = FORMC_QMAX;
    s.das.mu = d; d += FORMC_QMAX; s.das.r = d; d += FORMC_QMAX; s.das.y = d; d += FORMC_QMAX;
    s.das.wid = reinterpret_cast<int*>(d);
    s.das.wsg = reinterpret_cast<signed char*>(s.das.wid + FORMC_QMAX);
    s.das.state = s.das.wsg + FORMC_QMAX;
    s.das.qmax = FORMC_QMAX;
}

// MPCSolver.cpp:167-180: value of ftsp_midpoint(t, c).  rows: pointer to plan rows (x,y,z,t), first = index
// of rows[0] in the instance's plan.
__device__ __forceinline__ double midpoint_value(const double* rows, int first, int n_steps, int S, int F,
                                                 int t, int c)
{
    const int per = S + F;
    const int i = t / per, r = t - i * per;
    if (i >= n_steps - 1) return 0.0;                 // last step's rows stay 0 (loop runs to rows()-1)
    const double a = rows[(i - first) * 4 + c];
    if (r < S) return a * 1.0;
    const double b = rows[(i + 1 - first) * 4 + c];
    return a * 1.0 + (b - a) * ((double)(r - S) / (double)F);
}

// out_k = scale * sum_{j<k} (k-j) x_j  (the S_bar_z pattern: a prefix sum of a prefix sum, shifted by one), by one warp.
// Up to 4 elements per lane stay in registers (N <= 128); longer vectors go through the shared-memory scans.
__device__ __forceinline__ void warp_double_prefix_shifted(const double* x, double* out, double* scr, int N, double scale)
{
    const int lane = lane_id();
    int lo, hi; lane_chunk(N, lane, lo, hi);
    if (hi - lo <= 4 && ((N + 31) >> 5) <= 4) {
        double v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = lo + e < hi ? x[lo + e] : 0.0;
        v[1] += v[0]; v[2] += v[1]; v[3] += v[2];                       // first-level local prefix
        double incl = warp_incl_scan(v[3]);
        const double off1 = incl - v[3];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += off1;                      // P1_i for i in chunk (padding repeats the last)
        double w[4];
        w[0] = lo < hi ? v[0] : 0.0;
#pragma unroll
        for (int e = 1; e < 4; ++e) w[e] = w[e - 1] + (lo + e < hi ? v[e] : 0.0);
        const double t2 = w[3];                                          // padding adds 0: w[3] is the chunk total
        double incl2 = warp_incl_scan(t2);
        const double off2 = incl2 - t2;
        // out_k = scale * P2_{k-1}: lane writes out[i+1] for its i, out[0] = 0
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 4; ++e) if (lo + e < hi && lo + e + 1 < N) out[lo + e + 1] = scale * (w[e] + off2);
        if (lane == 0) out[0] = 0.0;
        __syncwarp();
        return;
    }
    for (int i = lane; i < N; i += 32) scr[i] = x[i];
    __syncwarp();
    warp_prefix_sum_smem(scr, N);
    warp_prefix_sum_smem(scr, N);
    for (int i = lane; i < N; i += 32) out[i] = (i == 0) ? 0.0 : scale * scr[i - 1];
    __syncwarp();
}

// ---- vertical QP policy for the dual active-set engine -------------------------------------------
struct VertProb {
    int N;
    FormCTables T;
    double c1;       // dt*dt/mass  (S_bar_z[k][j] = (k-j)*c1, MPCSolver.cpp:149)
    double fzmax;
    double* scr;     // [N] shared scratch
    __device__ int m() const { return N; }
    __device__ int nvar() const { return N; }
    __device__ double lo(int) const { return 0.0; }
    __device__ double hi(int) const { return fzmax; }
    __device__ void on_step(double) const {}
    // rv_k = (S_bar_z x)_k = c1 * sum_{j<k} (k-j) x_j  -> two prefix sums
    __device__ void eval(const double* x, double* rv) const { warp_double_prefix_shifted(x, rv, scr, N, c1); }
    __device__ double schur(int a, int b) const
    {
        if (a < N) return (b < N) ? T.M[(size_t)a * N + b] : T.G[(size_t)a * N + (b - N)];
        return (b < N) ? T.G[(size_t)b * N + (a - N)] : T.Hinv[(size_t)(a - N) * N + (b - N)];
    }
    __device__ const double* col(int id) const { return id < N ? T.G + (size_t)id * N : T.Hinv + (size_t)(id - N) * N; }
    __device__ void step_dir(int idp, int sgp, const int* wid, const signed char* wsg, const double* r, int q,
                             double* z) const
    {
        const int lane = lane_id();
        const double* cp = col(idp);
        for (int i = lane; i < N; i += 32) {
            double acc = (double)sgp * cp[i];
            for (int k = 0; k < q; ++k) acc -= (double)wsg[k] * r[k] * col(wid[k])[i];
            z[i] = acc;
        }
        __syncwarp();
    }
};

// Exact solve of   min 1/2|u|^2 - mid'u   s.t.  a'u = b,  mid-rho <= u <= mid+rho   by one warp.
// u_i = mid_i + clip(nu*a_i, -rho, rho); the equality residual is an odd, monotone, piecewise-linear function of nu
// (the scalar-Schur-complement view of the active-set method for this structure: H = I, A = [a'; I]).
//  * Direct path: the stability row decays along the horizon (with exact zeros on flight-phase ticks, where B = 0),
//    so the saturated rows are a PREFIX [0,k).  With P1 = prefix sums of |a| and S2 = suffix sums of a^2 every k has
//    the closed-form candidate t_k = (|r| - rho P1(k)) / S2(k); the one whose own saturation pattern is exactly that
//    prefix is the solution (checked, not assumed).
//    One pass, independent of how many rows saturate (a Newton pass per saturated row otherwise: the slowest QP of
//    a batch used to set the batch tick time).
//  * Fallback (no candidate is self-consistent, i.e. the saturated set is not a prefix): Newton on the multiplier; every pass adds all newly saturated rows, |nu| grows
//    monotonically, so it terminates in <= N passes.
// s1, s2: two scratch vectors [N] in shared memory owned by this warp.
// Returns status (0 ok, 1 infeasible), *nu_out, *iters (number of saturated rows / Newton passes - 1).
__device__ inline int knapsack_qp(int N, const double* a, const double* mid, double rho, double b, double* s1, double* s2,
                                  double* nu_out, int* iters_out, double* resid_out)
{
    const int lane = lane_id();
    double am = 0.0, aa = 0.0;
    for (int i = lane; i < N; i += 32) { am += a[i] * mid[i]; aa += a[i] * a[i]; }
    am = warp_sum(am); aa = warp_sum(aa);
    const double r = b - am;
    const double sg = (r >= 0.0) ? 1.0 : -1.0;
    const double rabs = fabs(r);
    int status = 0, iters = 0;
    double t = 0.0;           // |nu|
    if (aa > 0.0) {
        bool solved = false;
        {
            // Candidate k: rows [0,k) saturated (those with a_i = 0 never are and add nothing to either sum).
            //   P1(k) = sum_{i<k} |a_i|,  S2(k) = sum_{i>=k} a_i^2,  t_k = (|r| - rho P1(k)) / S2(k)
            // is THE solution iff its own saturation pattern is that prefix:
            //   t_k * min{|a_i| : i < k, a_i != 0} > rho   and   t_k * max{|a_i| : i >= k} <= rho.
            // Prefix quantities are laid down per k in an ascending sweep of the lane's chunk (s1, s2), suffix
            // quantities accumulate in the descending sweep that tests the candidates (small terms first).
            int lo, hi; lane_chunk(N, lane, lo, hi);
            const double INF = 1e300;
            double t1 = 0.0, t2 = 0.0, mn = INF, mx = 0.0;
            for (int i = lo; i < hi; ++i) {
                const double ai = fabs(a[i]);
                t1 += ai; t2 += ai * ai; mx = fmax(mx, ai);
                if (ai > 0.0) mn = fmin(mn, ai);
            }
            double p1 = t1, sf = t2, pm = mn, sm_ = mx;      // inclusive scans across lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double a1 = __shfl_up_sync(ISMPC_FULL_MASK, p1, o), a3 = __shfl_up_sync(ISMPC_FULL_MASK, pm, o);
                const double a2 = __shfl_down_sync(ISMPC_FULL_MASK, sf, o), a4 = __shfl_down_sync(ISMPC_FULL_MASK, sm_, o);
                if (lane >= o) { p1 += a1; pm = fmin(pm, a3); }
                if (lane + o < 32) { sf += a2; sm_ = fmax(sm_, a4); }
            }
            double P1 = p1 - t1;                                                       // exclusive prefix sum
            double Lm = __shfl_up_sync(ISMPC_FULL_MASK, pm, 1); if (lane == 0) Lm = INF;    // exclusive prefix min
            double S2 = sf - t2;                                                       // exclusive suffix sum
            double Mx = __shfl_down_sync(ISMPC_FULL_MASK, sm_, 1); if (lane == 31) Mx = 0.0; // exclusive suffix max
            for (int k = lo; k < hi; ++k) {
                s1[k] = P1; s2[k] = Lm;
                const double ak = fabs(a[k]);
                P1 += ak; if (ak > 0.0) Lm = fmin(Lm, ak);
            }
            int kbest = 0x7fffffff; double tbest = 0.0;
            for (int k = hi - 1; k >= lo; --k) {
                const double ak = fabs(a[k]);
                S2 += ak * ak; Mx = fmax(Mx, ak);
                if (!(S2 > 0.0)) continue;
                const double tk = (rabs - rho * s1[k]) / S2;
                const double lm = s2[k];
                const bool left = !(lm < INF) || (tk * lm > rho);
                const bool right = !(tk * Mx > rho);
                if (left && right && tk >= 0.0) { kbest = k; tbest = tk; }
            }
            double key = (double)kbest;
            int src = lane;
            warp_argmin(key, src);                // smallest valid k; src = the lane that holds it
            tbest = __shfl_sync(ISMPC_FULL_MASK, tbest, src);
            if (key < 2.0e9) {
                t = tbest; solved = true;
                int ns = 0;                       // report the number of saturated rows
                for (int i = lane; i < N; i += 32) ns += (t * fabs(a[i]) > rho);
                iters = warp_sum_int(ns) + 1;
            }
        }
        if (!solved) {
            int nsat_prev = -1;
            iters = 0;
            t = rabs / aa;
            for (;;) {
                double q1 = 0.0, q2 = 0.0; int ns = 0;
                for (int i = lane; i < N; i += 32) {
                    double ai = fabs(a[i]);
                    if (t * ai > rho) { q1 += ai; ++ns; } else q2 += ai * ai;
                }
                q1 = warp_sum(q1); q2 = warp_sum(q2); ns = warp_sum_int(ns);
                if (ns == nsat_prev) break;
                nsat_prev = ns; ++iters;
                double rem = rabs - rho * q1;
                if (!(q2 > 0.0)) { if (rem > 1e-12 * fmax(1.0, rabs)) status = 1; break; }
                double tn = rem / q2;
                if (!(tn >= t)) tn = t;   // monotone guard against round-off
                t = tn;
                if (iters > N + 2) break;
            }
        }
    } else if (rabs > 1e-12) status = 1;
    const double nu = sg * t;
    // equality residual of the final point (self-check)
    double au = 0.0;
    for (int i = lane; i < N; i += 32) au += a[i] * (mid[i] + fmin(fmax(nu * a[i], -rho), rho));
    au = warp_sum(au);
    *nu_out = nu; *iters_out = iters > 0 ? iters - 1 : 0; *resid_out = fabs(au - b);
    return status;
}

struct FormCArgs {
    int n;
    ismpc_formc_model_t model;
    FormCTables T;
    const ismpc_state_t* state;
    const ismpc_walk_t* walk;
    const ismpc_formc_inst_t* inst;
    const ismpc_formc_tick_t* tick;   // non-null (warp kernel family only): state / walk come packed, one 128-byte record per instance
    const double* plan;
    int plan_rows;
    ismpc_formc_out_t* out;
    double* primal;      // nullable, n x 3N
    signed char* active; // nullable, n x 3N
};

// Debug-only phase timing (make dbg -> lib/libismpc_b200_dbg.so; never in the product library).
#ifdef ISMPC_PHASE_TIMING
#define ISMPC_PHASE(k) do { if (threadIdx.x == 0 && blockIdx.x == (gridDim.x > 8 ? 5 : 0)) g_phase[k] = clock64(); } while (0)
#else
#define ISMPC_PHASE(k) do { } while (0)
#endif

// One tick for one instance by one CTA (FORMC_THREADS threads).  `st`/`wk` are the instance's current
// state/walk (registers, uniform across the CTA); results are written to *o (thread 0) and, if non-null,
// prim (3N) / act (3N).
// CLUSTER: the instance is owned by a thread-block cluster (long horizons).  The one O(N^2) stage -- the table
// mat-vec x0 = -H_z^-1 F_z, N^2 FMAs and N^2 doubles out of L2 -- is split by output rows over the CTAs of the
// cluster; every CTA then stores its slice of x0 into the shared memory of all its peers (distributed shared
// memory) and, after one cluster barrier, all CTAs carry on redundantly with the O(N) stages, so no further
// exchange is needed.  Only rank 0 writes results (the caller passes o/prim/act as null on the other ranks).
template <bool CLUSTER>
__device__ inline void formc_tick(const FormCShared& sm, const ismpc_formc_model_t& mdl, const FormCTables& T,
                                  const ismpc_state_t& st, const ismpc_walk_t& wk, const ismpc_formc_inst_t& in,
                                  const double* __restrict__ plan_all, ismpc_formc_out_t* o,
                                  double* prim, signed char* act, uint32_t& bar_parity)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = mdl.N;
    const double dt = mdl.dt, mass = mdl.mass, g = mdl.g;
    const double h = in.com_height;
    const double eta = sqrt(g / h);                       // parameters.cpp:41
    const int S = in.S, F = in.F_ds, per = S + F;
    const int k0 = (int)(wk.sim_time / (dt / mdl.dtc));   // MPCSolver.cpp:259,329
    int status = 0;

    if (k0 < 0 || per <= 0 || (long long)k0 + 2 * N > (long long)in.n_steps * per) {
        if (tid == 0 && o) {
            o->next = st; o->zmp_in[0] = o->zmp_in[1] = 0.0; o->fz0 = 0.0; o->lambda0 = 0.0; o->kkt_res = 0.0;
            o->status = ISMPC_ST_WINDOW; o->iters[0] = o->iters[1] = o->iters[2] = 0;
        }
        return;
    }

    // ---- stage the plan rows the window touches: one 1-D bulk (TMA) copy into shared memory ----------
    const int first = k0 / per;
    int last = (k0 + 2 * N - 1) / per + 1;                // inclusive: row i+1 is read for the ramp
    if (last > in.n_steps - 1) last = in.n_steps - 1;
    const int nrows = last - first + 1;
    const double* rows_g = plan_all + ((size_t)in.plan_first_row + first) * 4;
    const double* rows;
    if (nrows <= FORMC_PLAN_STAGE) {
        if (tid == 0) {
            mbar_expect_tx(sm.bar, (uint32_t)(nrows * 32));
            tma_load_1d(sm.plan, rows_g, (uint32_t)(nrows * 32), sm.bar);
        }
        mbar_wait(sm.bar, bar_parity);
        bar_parity ^= 1u;
        rows = sm.plan;
    } else {
        rows = rows_g;                                     // very short steps: read through L1/L2 instead
    }
    for (int k = tid; k < 2 * N; k += FORMC_THREADS) {
        sm.midx[k] = midpoint_value(rows, first, in.n_steps, S, F, k0 + k, 0);
        sm.midy[k] = midpoint_value(rows, first, in.n_steps, S, F, k0 + k, 1);
        if (k < N) sm.midz[k] = midpoint_value(rows, first, in.n_steps, S, F, k0 + k, 2);
    }
    __syncthreads();
    ISMPC_PHASE(0);

    // ================= STAGE 1: vertical QP (MPCSolver.cpp:220-269) =================
    const double z0 = st.com_pos[2], zd0 = st.com_vel[2];
    const double c1 = dt * dt / mass;                      // S_bar_z[k][j]   = (k-j)*dt * dt/m
    const double c1v = dt / mass;                          // S_bar_z_v[k][j] = dt/m          (j<k)
    // v_k = T_z[k] z + T_g[k] - h - mid_z[k];  w_k = T_zv[k] z + T_gv[k]           (:259)
    //   T_z[k] = [1, (k+1)dt], T_g[k] = -g dt^2 k(k+1)/2, T_zv[k] = [0,1], T_gv[k] = -g dt k
    for (int k = tid; k < N; k += FORMC_THREADS) {
        double tg = -g * (dt * dt) * (0.5 * (double)k * (double)(k + 1));
        sm.rv[k] = 1.0 * z0 + ((double)(k + 1) * dt) * zd0 + tg - h - sm.midz[k];   // v
        sm.scr[k] = zd0 - g * dt * (double)k;                                       // w
    }
    __syncthreads();
    ISMPC_PHASE(1);
    // F_j = q_p*c1 * sum_{k>j}(k-j) v_k + q_v*c1v * sum_{k>j} w_k - q_u*m*g
    if (warp == 0) {
        warp_suffix_sum_smem(sm.rv, N);    // SI_l = sum_{k>=l} v_k
        warp_suffix_sum_smem(sm.rv, N);    // sum_{l>=j} SI_l  -> P2_j = that at j+1
    } else if (warp == 1) {
        warp_suffix_sum_smem(sm.scr, N);   // sum_{k>=j} w_k
    }
    __syncthreads();
    ISMPC_PHASE(2);
    for (int j = tid; j < N; j += FORMC_THREADS) {
        double p2 = (j + 1 < N) ? sm.rv[j + 1] : 0.0;
        double pw = (j + 1 < N) ? sm.scr[j + 1] : 0.0;
        sm.Fz[j] = mdl.q_p * c1 * p2 + mdl.q_v * c1v * pw - mdl.q_u * mass * g;
    }
    __syncthreads();
    ISMPC_PHASE(3);
    // equalities f_k = 0 on the flight-phase columns (:223-243), active only when running (:262-269)
    int ne = 0, c_lo = 0;
    if (wk.footstep_counter > 1) {
        formc_flight_range(N, S, F, wk.mpc_iter, c_lo, ne);
    }
    // Prepared gait: the projector table of this mpcIter folds the equalities into the mat-vec.
    const bool use_P = T.P != nullptr && ne > 0 && S == T.gS && F == T.gF && wk.mpc_iter >= 0 && wk.mpc_iter < S + F;
    const double* __restrict__ Tab = use_P ? T.P + (size_t)wk.mpc_iter * N * N : T.Hinv;
    // minimiser f = -Tab F  (table mat-vec; the tables are symmetric -> coalesced row reads).
    // The 4 warps split the j range; each lane owns outputs i = blk*128 + lane + 32e.  Per-warp partial
    // vectors go to shared memory (zdir/lam/chv/shs are free at this point) and are summed afterwards.
    int i_lo = 0, i_hi = N;                                 // output rows of x0 this CTA computes
    if constexpr (CLUSTER) {
        const unsigned cr = cooperative_groups::this_cluster().block_rank(), cs = cooperative_groups::this_cluster().num_blocks();
        i_lo = (int)((long long)N * cr / cs); i_hi = (int)((long long)N * (cr + 1) / cs);
    }
    {
        double* part = (warp == 0) ? sm.zdir : (warp == 1) ? sm.lam : (warp == 2) ? sm.chv : sm.shs;
        const int jlo = (N * warp) / 4, jhi = (N * (warp + 1)) / 4;
        constexpr int U = 4;                                     // table rows in flight per pass (16 loads per lane)
        for (int blk = 0; i_lo + blk * 128 < i_hi; ++blk) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const int ib = i_lo + blk * 128 + lane;
            bool ok[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) ok[e] = ib + 32 * e < i_hi;
            int j = jlo;
            for (; j + U <= jhi; j += U) {
                double v[U][4], fj[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double* hr = Tab + (size_t)(j + u) * N + ib;
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[u][e] = ok[e] ? __ldg(hr + 32 * e) : 0.0;
                    fj[u] = sm.Fz[j + u];
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] += v[u][e] * fj[u];
            }
            for (; j < jhi; ++j) {
                const double fj = sm.Fz[j];
                const double* hr = Tab + (size_t)j * N + ib;
#pragma unroll
                for (int e = 0; e < 4; ++e) if (ok[e]) acc[e] += __ldg(hr + 32 * e) * fj;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) if (ok[e]) part[ib + 32 * e] = acc[e];
        }
    }
    __syncthreads();
    ISMPC_PHASE(4);
    for (int i = i_lo + tid; i < i_hi; i += FORMC_THREADS) {
        const double v = -(sm.zdir[i] + sm.lam[i] + sm.chv[i] + sm.shs[i]);
        sm.f[i] = v;
        if constexpr (CLUSTER) {
            auto cl = cooperative_groups::this_cluster();
            const unsigned cr = cl.block_rank(), cs = cl.num_blocks();
            for (unsigned p = 0; p < cs; ++p) if (p != cr) cl.map_shared_rank(sm.f, p)[i] = v;
        }
    }
    if constexpr (CLUSTER) cooperative_groups::this_cluster().sync(); else __syncthreads();
    ISMPC_PHASE(5);

    // keep the unconstrained minimiser for the general path (with a projector table sm.f is already the
    // equality-constrained minimiser; the general path then recomputes x0 from H^-1, see below)
    if (!use_P) for (int i = tid; i < N; i += FORMC_THREADS) sm.Fz[i] = sm.f[i];
    // ---- fast path: equality-constrained minimiser in closed form (all rows of 0 <= S f <= fz_max inactive) ----
    //   f = x0 - H^-1[:,K] mu,  (H^-1)_KK mu = x0_K   with K = [c_lo, c_lo+ne) a contiguous column range
    const bool fast_eq = (!use_P && ne > 0 && ne <= 32);
    if (fast_eq) {
        double* Sk = sm.das.Js;                                      // ne x ne, row-major (fits: 32*32 <= packed qmax)
        for (int e = tid; e < ne * ne; e += FORMC_THREADS) {
            int r = e / ne, c = e - r * ne;
            Sk[e] = __ldg(T.Hinv + (size_t)(c_lo + r) * N + c_lo + c);
        }
        __syncthreads();
        ISMPC_PHASE(6);
        if (warp == 0) {
            // LDL' (right-looking, lane = row, one reciprocal per column, no sqrt), then the two
            // triangular solves with the stored reciprocals; ne <= 32.  Sk[i][j] (j<i) := L_ij, Sk[j][j] := 1/d_j
            for (int j = 0; j < ne; ++j) {
                const double rd = 1.0 / Sk[j * ne + j];
                double lij = 0.0;
                if (lane > j && lane < ne) lij = Sk[lane * ne + j];          // a_ij (pre-division)
                __syncwarp();
                if (lane > j && lane < ne) {
                    const double l = lij * rd;
                    for (int c = j + 1; c <= lane; ++c) Sk[lane * ne + c] -= l * Sk[c * ne + j];   // uses a_cj (undivided)
                }
                __syncwarp();
                if (lane > j && lane < ne) Sk[lane * ne + j] = lij * rd;
                if (lane == j) Sk[j * ne + j] = rd;
                __syncwarp();
            }
            double yv = (lane < ne) ? sm.f[c_lo + lane] : 0.0;               // forward: L y = rhs
            for (int k = 0; k < ne; ++k) {
                const double yk = __shfl_sync(ISMPC_FULL_MASK, yv, k);
                if (lane > k && lane < ne) yv -= Sk[lane * ne + k] * yk;
            }
            if (lane < ne) yv *= Sk[lane * ne + lane];                        // D^-1
            for (int k = ne - 1; k >= 0; --k) {                              // backward: L' mu = y
                const double mk = __shfl_sync(ISMPC_FULL_MASK, yv, k);
                if (lane < k) yv -= Sk[k * ne + lane] * mk;
            }
            if (lane < ne) sm.das.mu[lane] = yv;
        }
        __syncthreads();
        ISMPC_PHASE(7);
        for (int i = tid; i < N; i += FORMC_THREADS) {
            double acc = sm.f[i];
            for (int k = 0; k < ne; ++k) acc -= __ldg(T.Hinv + (size_t)(c_lo + k) * N + i) * sm.das.mu[k];
            sm.f[i] = acc;
        }
        __syncthreads();
        ISMPC_PHASE(8);
    }
    int it_z = 0;
    if (warp == 0) {
        DasWork w = sm.das;
        w.q = 0; w.neq = 0;
        for (int i = lane; i < N; i += 32) w.state[i] = 0;
        __syncwarp();
        VertProb vp{N, T, c1, mdl.fz_max, sm.scr};
        int zfail = 0;
        bool done = false;
        if (fast_eq || ne == 0 || use_P) {
            vp.eval(sm.f, sm.rv);
            double worst = 0.0;
            for (int i = lane; i < N; i += 32) {
                double v = sm.rv[i];
                worst = fmin(worst, fmin(v + 1e-10, (mdl.fz_max - v) + 1e-10 * (1.0 + fabs(mdl.fz_max))));
            }
            worst = warp_min(worst);
            done = !(worst < 0.0);
        }
        if (!done) {
            // general path: dual active set from the unconstrained minimiser (equalities first, never dropped)
            if (use_P) {                       // rare: sm.Fz still holds F_z; x0 = -H^-1 F_z by this warp alone
                for (int i = lane; i < N; i += 32) {
                    double acc = 0.0;
                    for (int j = 0; j < N; ++j) acc += __ldg(T.Hinv + (size_t)j * N + i) * sm.Fz[j];
                    sm.f[i] = -acc;
                }
            } else {
                for (int i = lane; i < N; i += 32) sm.f[i] = sm.Fz[i];
            }
            __syncwarp();
            for (int e = 0; e < ne; ++e) {
                int rc = das_add_equality(vp, w, sm.f, sm.zdir, N + c_lo + e, sm.f[c_lo + e], 0.0);
                if (rc < 0) zfail = 1;
                __syncwarp();
            }
            w.neq = w.q;
            int rc = das_solve(vp, w, sm.f, sm.rv, sm.zdir, 4 * N + 16, &it_z);
            if (rc != 0 || zfail) status |= ISMPC_ST_Z_FAIL;
        }
        // active set of the S_bar_z rows
        if (act) for (int i = lane; i < N; i += 32) act[i] = w.state[i];
        // primal residual for the self-check: rows within bounds (rv holds S x of the last eval)
        double viol = 0.0;
        for (int i = lane; i < N; i += 32) viol = fmax(viol, fmax(-sm.rv[i], sm.rv[i] - mdl.fz_max));
        viol = warp_max(viol);
        if (lane == 0) { sm.red[0] = viol; sm.red[1] = (double)status; sm.red[2] = (double)it_z; }
    }
    __syncthreads();
    ISMPC_PHASE(9);
    double kkt = fmax(0.0, sm.red[0]);
    status = (int)sm.red[1];
    it_z = (int)sm.red[2];
    if (prim) for (int i = tid; i < N; i += FORMC_THREADS) prim[i] = sm.f[i];

    // ================= STAGE 2: lambda sequence (MPCSolver.cpp:296-309) =================
    // z_pos = S_bar_z f + T_z z + T_g, lambda_j = (g + (f_j/m - g)) / z_pos_j.  S_bar_z f is what the last row
    // evaluation of stage 1 left in sm.rv (feasibility check of the fast path / final pass of the active set).
    ISMPC_PHASE(10);
    for (int j = tid; j < N; j += FORMC_THREADS) {
        double sf = sm.rv[j];
        double tg = -g * (dt * dt) * (0.5 * (double)j * (double)(j + 1));
        double zp = sf + 1.0 * z0 + ((double)(j + 1) * dt) * zd0 + tg;
        double zacc = (1.0 / mass) * sm.f[j] - g;
        double lam = (g + zacc) / zp;
        sm.lam[j] = lam;
        if (lam < 2.0) { sm.chv[j] = 1.0; sm.shs[j] = dt; sm.ssh[j] = 0.0; }       // integrator (:353-355)
        else {
            double s = sqrt(lam);
            const double ex = exp(s * dt), ei = 1.0 / ex;                           // cosh / sinh from one exp
            const double ch = 0.5 * (ex + ei), sh = 0.5 * (ex - ei);
            sm.chv[j] = ch; sm.shs[j] = sh / s; sm.ssh[j] = s * sh;                 // (:357-360)
        }
        sm.dl[j] = exp(-dt * eta * (double)j);                                      // deltas (:183-184)
    }
    __syncthreads();
    ISMPC_PHASE(11);
    const double fz0 = sm.f[0];
    const double lam0 = sm.lam[0];
    double nz0 = 1.0 * z0 + dt * zd0, nz1 = zd0 + (dt / mass) * fz0 - dt * g;      // (:274)
    if (isnan(nz0)) { nz0 = h; status |= ISMPC_ST_NAN_GUARD; }                      // (:277-278)
    if (isnan(nz1)) { nz1 = 0.0; status |= ISMPC_ST_NAN_GUARD; }

    // ================= STAGE 3: horizontal QPs (MPCSolver.cpp:322-398) =================
    double ux0 = 0.0, uy0 = 0.0;
    int it_x = 0, it_y = 0;
    if (lam0 > 2.0) {
        // stability row a_i = C_sc * A_{N-1}...A_{i+1} B_i with C_sc = [1, 1/eta] (:351-379): backward
        // row-vector recurrence c_{i-1} = c_i A_i, a_i = c_i B_i, evaluated as a chunked warp scan.
        if (warp == 0) {
            int lo, hi; lane_chunk(N, lane, lo, hi);
            // local product P = A_{hi-1} ... A_{lo}   (row-vector convention: c_{lo-1} = c_{hi-1} * P)
            double p00 = 1, p01 = 0, p10 = 0, p11 = 1;
            for (int i = hi - 1; i >= lo; --i) {
                double a00 = sm.chv[i], a01 = sm.shs[i], a10 = sm.ssh[i], a11 = sm.chv[i];
                double n00 = p00 * a00 + p01 * a10, n01 = p00 * a01 + p01 * a11;
                double n10 = p10 * a00 + p11 * a10, n11 = p10 * a01 + p11 * a11;
                p00 = n00; p01 = n01; p10 = n10; p11 = n11;
            }
            // exclusive suffix scan over lanes: T_L = P_31 * P_30 * ... * P_{L+1}
            // inclusive first: I_L = P_31 ... P_L  via  I_L = I_{L+o} * (own partial)   (Hillis-Steele, down-shuffles)
            double i00 = p00, i01 = p01, i10 = p10, i11 = p11;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double q00 = __shfl_down_sync(ISMPC_FULL_MASK, i00, o), q01 = __shfl_down_sync(ISMPC_FULL_MASK, i01, o);
                double q10 = __shfl_down_sync(ISMPC_FULL_MASK, i10, o), q11 = __shfl_down_sync(ISMPC_FULL_MASK, i11, o);
                if (lane + o < 32) {
                    double n00 = q00 * i00 + q01 * i10, n01 = q00 * i01 + q01 * i11;
                    double n10 = q10 * i00 + q11 * i10, n11 = q10 * i01 + q11 * i11;
                    i00 = n00; i01 = n01; i10 = n10; i11 = n11;
                }
            }
            // T_L = I_{L+1}; identity for lane 31
            double t00 = __shfl_down_sync(ISMPC_FULL_MASK, i00, 1), t01 = __shfl_down_sync(ISMPC_FULL_MASK, i01, 1);
            double t10 = __shfl_down_sync(ISMPC_FULL_MASK, i10, 1), t11 = __shfl_down_sync(ISMPC_FULL_MASK, i11, 1);
            if (lane == 31) { t00 = 1; t01 = 0; t10 = 0; t11 = 1; }
            const double cs0 = 1.0, cs1 = 1.0 / eta;                                 // C_sc (:375-377), nominal eta
            double c0 = cs0 * t00 + cs1 * t10, c1r = cs0 * t01 + cs1 * t11;          // c_{hi-1}
            for (int i = hi - 1; i >= lo; --i) {
                // B_i = [1-ch; -s*sh]; integrator branch has B = 0 (ssh = 0 and ch = 1)
                sm.avec[i] = c0 * (1.0 - sm.chv[i]) + c1r * (-sm.ssh[i]);
                double n0 = c0 * sm.chv[i] + c1r * sm.ssh[i], n1 = c0 * sm.shs[i] + c1r * sm.chv[i];
                c0 = n0; c1r = n1;
            }
            if (lane == 0) { sm.red[4] = c0; sm.red[5] = c1r; }                      // C_sc * phi_state
        } else if (warp == 1 || warp == 2) {
            const double* mq = (warp == 1) ? sm.midx : sm.midy;
            double t = 0.0;
            for (int i = lane; i < N; i += 32) t += sm.dl[i] * mq[N + i];            // anticipative tail (:381-383)
            t = warp_sum(t);
            if (lane == 0) sm.red[6 + (warp - 1)] = t;
        }
        __syncthreads();
        ISMPC_PHASE(12);
        if (warp < 2) {
            const int ax = warp;
            const double* mq = ax == 0 ? sm.midx : sm.midy;
            const double c = st.com_pos[ax], cd = st.com_vel[ax];
            const double b = -(sm.red[4] * c + sm.red[5] * cd) + eta * dt * sm.red[6 + ax];
            const double rho = (wk.footstep_counter > 1) ? in.box_w / 2 : in.box_w_init / 2;   // (:328-338)
            double nu, resid; int it;
            int rc = knapsack_qp(N, sm.avec, mq, rho, b, ax == 0 ? sm.zdir : sm.Fz, ax == 0 ? sm.scr : sm.rv, &nu, &it, &resid);
            for (int i = lane; i < N; i += 32) {
                double d = nu * sm.avec[i];
                double u = mq[i] + fmin(fmax(d, -rho), rho);
                if (prim) prim[(1 + ax) * N + i] = u;
                if (act) act[(1 + ax) * N + i] = (signed char)((d > rho) ? 1 : ((d < -rho) ? -1 : 0));
            }
            if (lane == 0) {
                double d0 = nu * sm.avec[0];
                sm.red[8 + ax * 3 + 0] = mq[0] + fmin(fmax(d0, -rho), rho);
                sm.red[8 + ax * 3 + 1] = (double)(rc ? (ax == 0 ? ISMPC_ST_X_FAIL : ISMPC_ST_Y_FAIL) : 0) + 1024.0 * it;
                sm.red[8 + ax * 3 + 2] = resid;
            }
        }
        __syncthreads();
        ISMPC_PHASE(13);
        ISMPC_PHASE(14);
        ux0 = sm.red[8]; uy0 = sm.red[11];
        int e0 = (int)sm.red[9], e1 = (int)sm.red[12];
        status |= (e0 & 1023) | (e1 & 1023);
        it_x = e0 >> 10; it_y = e1 >> 10;
        kkt = fmax(kkt, fmax(sm.red[10], sm.red[13]));
    } else {
        status |= ISMPC_ST_XY_SKIPPED;
        if (prim) for (int i = tid; i < 2 * N; i += FORMC_THREADS) prim[N + i] = 0.0;
        if (act) for (int i = tid; i < 2 * N; i += FORMC_THREADS) act[N + i] = 0;
    }

    // ================= integrate (MPCSolver.cpp:402-422) =================
    if (tid == 0 && o) {
        double a00, a01, a10, a11, b0, b1;
        if (lam0 < 2.0) { a00 = 1.0; a01 = dt; a10 = 0.0; a11 = 1.0; b0 = 0.0; b1 = 0.0; }
        else {
            double s = sqrt(lam0);
            const double ex = exp(s * dt), ei = 1.0 / ex;
            const double ch = 0.5 * (ex + ei), sh = 0.5 * (ex - ei);
            a00 = ch; a01 = sh / s; a10 = s * sh; a11 = ch; b0 = 1.0 - ch; b1 = -s * sh;
        }
        ismpc_formc_out_t r;
        r.next = st;
        r.next.com_pos[0] = a00 * st.com_pos[0] + a01 * st.com_vel[0] + b0 * ux0;
        r.next.com_vel[0] = a10 * st.com_pos[0] + a11 * st.com_vel[0] + b1 * ux0;
        r.next.com_pos[1] = a00 * st.com_pos[1] + a01 * st.com_vel[1] + b0 * uy0;
        r.next.com_vel[1] = a10 * st.com_pos[1] + a11 * st.com_vel[1] + b1 * uy0;
        r.next.com_pos[2] = nz0; r.next.com_vel[2] = nz1;
        r.zmp_in[0] = ux0; r.zmp_in[1] = uy0; r.fz0 = fz0; r.lambda0 = lam0; r.kkt_res = kkt;
        r.status = status; r.iters[0] = it_z; r.iters[1] = it_x; r.iters[2] = it_y;
        *o = r;
    }
}

}  // namespace ismpc
