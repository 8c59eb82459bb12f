// forma.cuh -- formulation A (canonical ISMPC with footsteps, the MATLAB scripts' first QP).
//
// Reference map (trotting/quad_as_bip_bang.m == trotting/quad_as_bip_no_plots.m == walking/quad_walk_no_plots.m):
//   bang.m:126-140  mapping tick -> footstep weights          -> forma_mapping()
//   bang.m:142-150  ZMP rows (Pzmp = tril(1)*dt, -mapping)     -> FormAProb::eval / schur / step_dir (never formed)
//   bang.m:156-190  kinematic rows                             -> same
//   bang.m:195-210  stability rows + anticipative tail         -> forma_stability()
//   bang.m:239-245  H = blkdiag(I_C, Qf I_F) x2, f             -> diagonal H^-1 inside FormAProb
//   bang.m:256      quadprog(...)                              -> das_solve() (dual active set, das.cuh)
//   bang.m:265-290  LIP 3-state update                         -> forma_integrate()
//   bang.m:529-563  footstep switch, plan shift, centerline    -> forma_rollout_kernel
//
// x and y are separable in this QP (H block diagonal, every row touches one axis), so one warp solves one
// (instance, axis) pair: variables v = [zd(C); xf(F)], 1 equality, C two-sided ZMP rows, F two-sided kin rows.
#pragma once
#include "common.cuh"
#include "das.cuh"
#include "../../include/ismpc_b200.h"

namespace ismpc {

struct FormAArgs {
    int n;
    ismpc_forma_model_t model;
    const ismpc_forma_inst_t* inst;
    const int32_t* fs_timing;
    int timing_len;
    const double* fs_plan;
    int plan_rows;
    ismpc_forma_out_t* out;
    double* primal;        // nullable, n x 2(C+F)
    signed char* active;   // nullable, n x 2(C+F)
    int sm_count;
    int warps_per_cta;
    int R;                 // rows of the inverse factor kept in shared memory (das.cuh)
    double* Jspill;        // global slices for rows >= R: one per resident warp (grid * warps_per_cta)
    int* queue;            // work queue head (items = (instance, axis) pairs), zeroed before the launch
};

struct FormAShared {   // per-warp slices
    double *a, *PA, *mw, *scr;     // [C]
    double *x, *z, *rv, *lo, *hi;  // [C+F]
    signed char* mp;               // [C]
    DasWork das;
};

__host__ __device__ inline size_t forma_vec_doubles(int C, int F) { return (size_t)4 * C + 5 * (C + F) + 3 * (C + F + 1); }
__host__ __device__ inline size_t forma_spill_doubles(int C, int F, int R)
{
    const int q = C + F + 1;
    return R >= q ? 0 : (size_t)(tri(q, 0) - tri(R, 0));
}
__host__ __device__ inline size_t forma_byte_tail(int C, int F)
{
    size_t q = C + F + 1;
    size_t b = q * sizeof(int) + q /*wsg*/ + (C + F) /*state*/ + C /*mp*/;
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t forma_warp_smem_bytes(int C, int F, int R)
{
    return (forma_vec_doubles(C, F) + (size_t)tri(R, 0)) * sizeof(double) + forma_byte_tail(C, F);
}

__device__ inline void forma_carve(unsigned char* base, int C, int F, int R, double* J_global, FormAShared& s)
{
    const int n = C + F, q = n + 1;
    double* d = reinterpret_cast<double*>(base);
    s.a = d; d += C; s.PA = d; d += C; s.mw = d; d += C; s.scr = d; d += C;
    s.x = d; d += n; s.z = d; d += n; s.rv = d; d += n; s.lo = d; d += n; s.hi = d; d += n;
    s.das.mu = d; d += q; s.das.r = d; d += q; s.das.y = d; d += q;
    s.das.Js = d; d += tri(R, 0);
    s.das.Jg = J_global; s.das.R = R;
    s.das.wid = reinterpret_cast<int*>(d);
    s.das.wsg = reinterpret_cast<signed char*>(s.das.wid + q);
    s.das.state = s.das.wsg + q;
    s.mp = s.das.state + n;
    s.das.qmax = q;
}

// quad_as_bip_bang.m:74-84 (initial) / :547-555 (rebuilt): centerline sample cl(t), t 1-based.
__device__ __forceinline__ double forma_centerline(const double* plan /*rows x 2*/, int n_fs, int axis, int step,
                                                   int ds, int first_ramp, int t)
{
    int seg = (t - 1) / step, r = (t - 1) - seg * step;
    if (seg > n_fs - 2) { seg = n_fs - 2; r = step - 1; }
    const double a = plan[seg * 2 + axis], b = plan[(seg + 1) * 2 + axis];
    if (seg == 0 && !first_ramp) return a;
    if (r < step - ds) return a;
    const int k = r - (step - ds);
    if (ds == 1 || k == ds - 1) return b;
    return a + (double)k * ((b - a) / (double)(ds - 1));
}

struct FormAProb {
    int C, F;
    double dt, qz_inv, qf_inv, saa;
    const double *a, *PA, *mw, *lo_, *hi_;
    const signed char* mp;
    double* scr;
    __device__ int m() const { return C + F; }
    __device__ int nvar() const { return C + F; }
    __device__ double lo(int i) const { return lo_[i]; }
    __device__ double hi(int i) const { return hi_[i]; }
    __device__ void on_step(double) const {}
    // coefficient of footstep variable f (0-based, column f+1 of `mapping`) in ZMP row i
    __device__ double mcoef(int i, int f) const
    {
        const int p = mp[i];
        const double w = mw[i];
        return (p == f + 1 ? w : 0.0) + (p + 1 == f + 1 ? 1.0 - w : 0.0);
    }
    __device__ void eval(const double* x, double* rv) const
    {
        const int lane = lane_id();
        for (int i = lane; i < C; i += 32) scr[i] = x[i];
        __syncwarp();
        warp_prefix_sum_smem(scr, C);
        for (int i = lane; i < C; i += 32) {
            double s = dt * scr[i];
            const int p = mp[i];
            const double w = mw[i];
            if (p >= 1) s -= w * x[C + p - 1];
            if (p + 1 <= F) s -= (1.0 - w) * x[C + p];
            rv[i] = s;
        }
        for (int f = lane; f < F; f += 32) rv[C + f] = x[C + f] - (f > 0 ? x[C + f - 1] : 0.0);
        __syncwarp();
    }
    __device__ double schur(int ia, int ib) const
    {
        const int n = C + F;
        if (ia > ib) { int t = ia; ia = ib; ib = t; }    // ia <= ib; order: ZMP rows < kin rows < equality
        if (ib == n) {                                    // with the stability row
            if (ia == n) return saa * qz_inv;
            if (ia < C) return dt * qz_inv * PA[ia];
            return 0.0;
        }
        if (ib < C) {                                     // ZMP-ZMP
            double s = dt * dt * qz_inv * (double)(ia + 1);
            double mm = 0.0;
            for (int f = 0; f < F; ++f) mm += mcoef(ia, f) * mcoef(ib, f);
            return s + qf_inv * mm;
        }
        const int g = ib - C;
        if (ia < C) return qf_inv * (-mcoef(ia, g) + (g > 0 ? mcoef(ia, g - 1) : 0.0));   // ZMP-kin
        const int f = ia - C;                              // kin-kin, f <= g
        if (f == g) return qf_inv * (f > 0 ? 2.0 : 1.0);
        if (g - f == 1) return -qf_inv;
        return 0.0;
    }
    __device__ void step_dir(int idp, int sgp, const int* wid, const signed char* wsg, const double* r, int q,
                             double* z) const
    {
        const int lane = lane_id();
        const int n = C + F;
        for (int i = lane; i < C; i += 32) scr[i] = 0.0;
        __syncwarp();
        double vf[ISMPC_MAX_FSTEPS];
#pragma unroll
        for (int f = 0; f < ISMPC_MAX_FSTEPS; ++f) vf[f] = 0.0;
        double ceq = 0.0;
        // entries of W plus the entering constraint (index q)
        for (int k = lane; k <= q; k += 32) {
            const int id = (k == q) ? idp : wid[k];
            const double coef = (k == q) ? (double)sgp : -(double)wsg[k] * r[k];
            if (id < C) {
                scr[id] = coef;                            // distinct ids -> distinct addresses
                const int p = mp[id];
                const double w = mw[id];
                if (p >= 1) vf[p - 1] -= coef * w;
                if (p + 1 <= F) vf[p] -= coef * (1.0 - w);
            } else if (id < n) {
                const int f = id - C;
                vf[f] += coef;
                if (f > 0) vf[f - 1] -= coef;
            } else ceq += coef;
        }
        __syncwarp();
        ceq = warp_sum(ceq);
#pragma unroll
        for (int f = 0; f < ISMPC_MAX_FSTEPS; ++f) if (f < F) vf[f] = warp_sum(vf[f]);
        warp_suffix_sum_smem(scr, C);                      // scr[k] = sum of coefs of ZMP rows with index >= k
        for (int k = lane; k < C; k += 32) z[k] = (dt * scr[k] + ceq * a[k]) * qz_inv;
#pragma unroll
        for (int f = 0; f < ISMPC_MAX_FSTEPS; ++f) if (f < F && lane == 0) z[C + f] = vf[f] * qf_inv;
        __syncwarp();
    }
};

// One tick for one (instance, axis) by one warp.  Returns status bits; writes x (primal) in sm.x.
// in: the instance (by reference; read only), plan: this instance's fs_plan rows, ft: its fs_timing.
__device__ inline int forma_tick_axis(const FormAShared& sm, const ismpc_forma_model_t& mdl,
                                      const ismpc_forma_inst_t& in, const double* st3 /*x,xd,xz of this axis*/,
                                      double cur, double fs_store, int j, int fs_counter, int first_ramp,
                                      const double* plan, const int32_t* ft, int axis,
                                      int* iters_out, double* kkt_out)
{
    const int lane = lane_id();
    const int C = mdl.C, P = mdl.P, F = mdl.F, n = C + F;
    const double dt = mdl.dt;
    const double eta = sqrt(mdl.g_eta / in.height);               // bang.m:31
    const double w_box = axis == 0 ? in.wx : in.wy;
    const int ds = in.ds, n_timing = in.n_timing, n_fs = in.n_fs;
    const int step = ft[1] - ft[0];
    // ---- stability row coefficients (bang.m:200-207) ----
    const double lam = exp(-eta * dt);
    const double k1 = (1.0 / eta) * (1.0 - lam) / (1.0 - pow(lam, (double)C));
    const double k2 = dt * exp(-eta * dt * (double)C);
    double saa = 0.0;
    for (int i = lane; i < C; i += 32) {
        double ai = k1 * exp(-eta * dt * (double)i) - k2;
        sm.a[i] = ai; sm.PA[i] = ai; saa += ai * ai;
    }
    saa = warp_sum(saa);
    __syncwarp();
    warp_prefix_sum_smem(sm.PA, C);
    // ---- mapping (bang.m:126-140) + ZMP bounds (bang.m:147-150) ----
    const double zq = st3[2];
    for (int i = lane; i < C; i += 32) {
        const int t = j + i + 1;                                    // MATLAB j+i with i 1-based
        int pf = 0;
        for (int mstep = 1; mstep <= F + 1; ++mstep) {
            int idx = fs_counter + mstep;                           // 1-based fs_timing index
            if (idx <= n_timing && t >= ft[idx - 1]) pf = mstep; else break;
        }
        int idx = fs_counter + pf + 1;
        if (idx > n_timing) idx = n_timing;
        const int rem = ft[idx - 1] - t;
        double wgt = (rem > ds) ? 1.0 : (double)rem / (double)ds;
        sm.mw[i] = wgt; sm.mp[i] = (signed char)pf;
        const double m0 = (pf == 0) ? wgt : 0.0;                    // mapping(:,1): weight on the current footstep
        sm.lo[i] = 1.0 * (-zq - w_box / 2) + m0 * cur;
        sm.hi[i] = 1.0 * (-zq + w_box / 2) + m0 * cur;
    }
    // ---- kinematic bounds (bang.m:163-190) ----
    for (int f = lane; f < F; f += 32) {
        double bnd = (axis == 0) ? ((fs_counter == 1 && f == 0) ? mdl.disp_forw_dummy : mdl.disp_forw)
                                 : (mdl.disp_L / 2 + mdl.disp_L / 2);
        double c0 = (f == 0) ? cur : 0.0;
        sm.lo[C + f] = -bnd + c0; sm.hi[C + f] = bnd + c0;
    }
    // ---- anticipative tail (bang.m:195-198); cl(P) uses the ABSOLUTE index P, copied as written ----
    double ant = 0.0;
    for (int i = C + 1 + lane; i <= P; i += 32)
        ant += exp(-eta * dt * (double)i) * (1.0 - exp(-eta * dt)) *
               (forma_centerline(plan, n_fs, axis, step, ds, first_ramp, j + i) - fs_store);
    ant = warp_sum(ant);
    ant += exp(-eta * dt * (double)P) * (forma_centerline(plan, n_fs, axis, step, ds, first_ramp, P) - fs_store);
    const double beq = st3[0] + st3[1] / eta - st3[2] - ant;       // bang.m:209-210
    // ---- unconstrained minimiser: zd = 0, xf = plan(fsCounter+1 .. fsCounter+F) (bang.m:244-245) ----
    for (int i = lane; i < C; i += 32) sm.x[i] = 0.0;
    for (int f = lane; f < F; f += 32) {
        int row = fs_counter + 1 + f; if (row > n_fs) row = n_fs;
        sm.x[C + f] = plan[(row - 1) * 2 + axis];
    }
    for (int i = lane; i < n; i += 32) sm.das.state[i] = 0;
    __syncwarp();
    FormAProb pb{C, F, dt, 1.0 / mdl.q_zdot, 1.0 / mdl.q_foot, saa, sm.a, sm.PA, sm.mw, sm.lo, sm.hi, sm.mp, sm.scr};
    DasWork w = sm.das;
    w.q = 0; w.neq = 0;
    int status = 0;
    int rc = das_add_equality(pb, w, sm.x, sm.z, n, 0.0, beq);
    if (rc < 0) status |= ISMPC_ST_QP_FAIL;
    w.neq = w.q;
    int iters = 0;
    rc = das_solve(pb, w, sm.x, sm.rv, sm.z, 6 * n + 50, &iters);
    if (rc != 0) status |= ISMPC_ST_QP_FAIL;
    // self-check: equality residual and worst bound violation
    double eqv = 0.0;
    for (int i = lane; i < C; i += 32) eqv += sm.a[i] * sm.x[i];
    eqv = warp_sum(eqv);
    double viol = 0.0;
    for (int i = lane; i < n; i += 32) viol = fmax(viol, fmax(sm.lo[i] - sm.rv[i], sm.rv[i] - sm.hi[i]));
    viol = warp_max(viol);
    *iters_out = iters;
    *kkt_out = fmax(fabs(eqv - beq), fmax(viol, 0.0));
    return status;
}

// bang.m:55-58,265-290: [c; cd; z]+ = A_upd [c; cd; z] + B_upd * zd(1)
__device__ __forceinline__ void forma_integrate(double eta, double dt, double s[3], double zd0)
{
    const double ch = cosh(eta * dt), sh = sinh(eta * dt);
    const double n0 = ch * s[0] + (sh / eta) * s[1] + (1 - ch) * s[2] + (dt - sh / eta) * zd0;
    const double n1 = (eta * sh) * s[0] + ch * s[1] + (-eta * sh) * s[2] + (1 - ch) * zd0;
    const double n2 = 0 * s[0] + 0 * s[1] + 1 * s[2] + dt * zd0;
    s[0] = n0; s[1] = n1; s[2] = n2;
}

}  // namespace ismpc
