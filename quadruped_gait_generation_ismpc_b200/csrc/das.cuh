// das.cuh -- warp-per-QP dual active-set engine (Goldfarb-Idnani in Schur-complement form), FP64.
//
// Replaces the role qpOASES' QProblem::init/hotstart plays behind the reference's solveQP wrapper
// (AMR_code_DART/utils.cpp:121-130 -> qpOASES/QProblem.cpp:302,1279,432,1455) for strictly convex
// QPs whose Hessian inverse is cheap to apply (identity, diagonal, or a precomputed table).  It is a
// different algorithm (dual, not qpOASES' primal-dual homotopy): the QPs on this path are strictly
// convex, so the minimiser -- and with it the strongly active set -- is unique and both methods
// must agree.
//
// One warp owns one QP.  All state lives in shared memory slices owned by that warp; all control
// flow is warp-uniform; synchronisation is __syncwarp only.
//
// Working set W = {(id_k, sg_k)}, multipliers mu_k >= 0 (free sign for the first `neq` entries,
// which are equalities and are never dropped).  Normal of entry k: n_k = sg_k * a_{id_k}.
// S = N' H^-1 N (q x q) is kept as a packed lower Cholesky factor L; adding a constraint appends a
// row (one forward solve), dropping one deletes a row/column and repairs the trailing block with a
// rank-1 update.
//
// Problem policy P (all methods warp-collective, results warp-uniform):
//   int    m()                          number of two-sided inequality rows (ids 0..m-1; equality ids >= m)
//   double lo(i), hi(i)                 bounds of inequality row i
//   void   eval(const double* x, double* rv)          rv[i] = a_i' x for i < m   (shared memory out)
//   double schur(int ida, int idb)      a_ida' H^-1 a_idb  (called by single lanes: must be lane-local)
//   void   step_dir(int idp, int sgp, const int* wid, const signed char* wsg, const double* r, int q,
//                   double* z)          z = H^-1 (sgp*a_idp - sum_k wsg_k r_k a_{wid_k})  (shared memory out)
//   int    nvar()
#pragma once
#include "common.cuh"

namespace ismpc {

struct DasWork {
    double* L;          // packed lower triangular, row-major: L[i*(i+1)/2 + j], j <= i < qmax
    double* mu;         // [qmax]
    double* r;          // [qmax]
    double* y;          // [qmax]
    int* wid;           // [qmax]
    signed char* wsg;   // [qmax]
    signed char* state; // [m] 0 free, -1 lower active, +1 upper active
    int q, neq, qmax;
};

__device__ __forceinline__ int tri(int i, int j) { return ((i * (i + 1)) >> 1) + j; }

// y = L^-1 c (forward substitution, in place in w.y), returns y'y.  Column-oriented: lane owns rows i == lane (mod 32).
__device__ __forceinline__ double das_forward(const DasWork& w)
{
    const int lane = lane_id();
    const int q = w.q;
    double yy = 0.0;
    for (int k = 0; k < q; ++k) {
        double yk = w.y[k] / w.L[tri(k, k)];   // broadcast read (same address for all lanes)
        __syncwarp();
        if (lane == 0) w.y[k] = yk;
        yy += yk * yk;
        for (int i = k + 1 + lane; i < q; i += 32) w.y[i] -= w.L[tri(i, k)] * yk;
        __syncwarp();
    }
    return yy;
}
// r = L^-T y (backward substitution) into w.r.
__device__ __forceinline__ void das_backward(const DasWork& w)
{
    const int lane = lane_id();
    const int q = w.q;
    for (int i = lane; i < q; i += 32) w.r[i] = w.y[i];
    __syncwarp();
    for (int k = q - 1; k >= 0; --k) {
        double rk = w.r[k] / w.L[tri(k, k)];
        __syncwarp();
        if (lane == 0) w.r[k] = rk;
        for (int i = lane; i < k; i += 32) w.r[i] -= w.L[tri(k, i)] * rk;
        __syncwarp();
    }
}

// Delete entry l from the working set: shift bookkeeping, delete row/col l of S and repair L.
__device__ __forceinline__ void das_drop(DasWork& w, int l)
{
    const int lane = lane_id();
    const int q = w.q;
    // save column l below the diagonal into w.y (v), indexed by NEW row index
    for (int i = l + 1 + lane; i < q; i += 32) w.y[i - 1] = w.L[tri(i, l)];
    __syncwarp();
    // move rows up / columns left, row by row (targets of row i' never overlap sources of row i'+1)
    for (int ip = l; ip < q - 1; ++ip) {
        const int io = ip + 1;
        // elements j' in [0, ip]: from (io, j') if j' < l else (io, j'+1)
        double buf[4];
        int cnt = 0;
        for (int jp = lane; jp <= ip; jp += 32) {
            int jo = jp < l ? jp : jp + 1;
            buf[cnt & 3] = w.L[tri(io, jo)];
            // qmax <= 128 -> at most 4 elements per lane per row
            ++cnt;
        }
        __syncwarp();
        cnt = 0;
        for (int jp = lane; jp <= ip; jp += 32) { w.L[tri(ip, jp)] = buf[cnt & 3]; ++cnt; }
        __syncwarp();
    }
    // shift wid/wsg/mu serially by lane 0 (q is small; avoids read/write overlap hazards)
    if (lane == 0) {
        for (int k = l; k < q - 1; ++k) { w.wid[k] = w.wid[k + 1]; w.wsg[k] = w.wsg[k + 1]; w.mu[k] = w.mu[k + 1]; }
    }
    __syncwarp();
    w.q = q - 1;
    // rank-1 update of the trailing block (rows/cols >= l of the new factor) with v = w.y[l..q-2]
    const int qn = q - 1;
    for (int k = l; k < qn; ++k) {
        double lkk = w.L[tri(k, k)];
        double vk = w.y[k];
        double rr = sqrt(lkk * lkk + vk * vk);
        double c = rr / lkk, s = vk / lkk;
        __syncwarp();
        if (lane == 0) w.L[tri(k, k)] = rr;
        for (int i = k + 1 + lane; i < qn; i += 32) {
            double lik = w.L[tri(i, k)];
            double vi = w.y[i];
            double nl = (lik + s * vi) / c;
            w.L[tri(i, k)] = nl;
            w.y[i] = c * vi - s * nl;
        }
        __syncwarp();
    }
}

// Append constraint (id, sg) with multiplier mu0: new Cholesky row = [y, sqrt(zn)].
__device__ __forceinline__ void das_append(DasWork& w, int id, int sg, double zn, double mu0)
{
    const int lane = lane_id();
    const int q = w.q;
    for (int j = lane; j < q; j += 32) w.L[tri(q, j)] = w.y[j];
    if (lane == 0) {
        w.L[tri(q, q)] = sqrt(zn);
        w.wid[q] = id; w.wsg[q] = (signed char)sg; w.mu[q] = mu0;
    }
    __syncwarp();
    w.q = q + 1;
}

// Schur column c_k = wsg_k * sgp * a_{wid_k}' H^-1 a_p into w.y; returns spp = a_p' H^-1 a_p.
template <class P>
__device__ __forceinline__ double das_schur_col(const P& prob, const DasWork& w, int idp, int sgp)
{
    const int lane = lane_id();
    for (int k = lane; k < w.q; k += 32) w.y[k] = (double)(w.wsg[k] * sgp) * prob.schur(w.wid[k], idp);
    double spp = prob.schur(idp, idp);
    __syncwarp();
    return spp;
}

// Add an equality row (id >= m) to the working set, moving x onto it.  value = a_id' x (current), target = rhs.
// Returns 0 ok, 1 dependent-and-skipped, -1 inconsistent.
template <class P>
__device__ int das_add_equality(P& prob, DasWork& w, double* x, double* z, int id, double value, double target)
{
    const int lane = lane_id();
    double spp = das_schur_col(prob, w, id, +1);
    double yy = das_forward(w);
    double zn = spp - yy;
    if (!(zn > 1e-13 * spp)) return (fabs(value - target) <= 1e-9 * fmax(1.0, fabs(target))) ? 1 : -1;
    das_backward(w);
    double t = (target - value) / zn;
    prob.step_dir(id, +1, w.wid, w.wsg, w.r, w.q, z);
    const int n = prob.nvar();
    for (int i = lane; i < n; i += 32) x[i] += t * z[i];
    for (int k = lane; k < w.q; k += 32) w.mu[k] -= t * w.r[k];
    __syncwarp();
    if (w.q >= w.qmax) return -1;
    das_append(w, id, +1, zn, t);
    return 0;
}

// Main loop.  On entry x satisfies stationarity for the current (W, mu).  Returns 0 solved, 1 infeasible,
// 2 iteration cap, 3 working set overflow.  *iters_out = number of working-set changes.
template <class P>
__device__ int das_solve(P& prob, DasWork& w, double* x, double* rv, double* z, int maxit, int* iters_out)
{
    const int lane = lane_id();
    const int m = prob.m();
    const int n = prob.nvar();
    int iters = 0;
    int rc = 0;
    for (;;) {
        prob.eval(x, rv);
        // most violated inequality row not in W
        double best = 0.0; int bidx = 0x7fffffff;
        for (int i = lane; i < m; i += 32) {
            if (w.state[i] != 0) continue;
            double lo = prob.lo(i), hi = prob.hi(i), v = rv[i];
            double sl = v - lo, su = hi - v;
            double tl = 1e-10 * (1.0 + fabs(lo)), tu = 1e-10 * (1.0 + fabs(hi));
            if (sl < -tl && sl < best) { best = sl; bidx = 2 * i; }
            if (su < -tu && su < best) { best = su; bidx = 2 * i + 1; }
        }
        warp_argmin(best, bidx);
        if (bidx == 0x7fffffff) break;   // primal feasible -> optimal
        const int p = bidx >> 1;
        const int sgp = (bidx & 1) ? -1 : +1;
        double sviol = best;             // n_p' x - beta_p  (< 0)
        double up = 0.0;
        bool added = false;
        while (!added) {
            if (++iters > maxit) { rc = 2; goto done; }
            double spp = das_schur_col(prob, w, p, sgp);
            double yy = das_forward(w);
            double zn = spp - yy;
            das_backward(w);
            // dual ratio test over droppable entries
            double t1 = 1e300; int l = 0x7fffffff;
            for (int k = w.neq + lane; k < w.q; k += 32) {
                double rk = w.r[k];
                if (rk > 1e-14) { double t = w.mu[k] / rk; if (t < t1) { t1 = t; l = k; } }
            }
            warp_argmin(t1, l);
            const bool dependent = !(zn > 1e-13 * spp);
            double t2 = dependent ? 1e300 : fmax(0.0, -sviol / zn);
            double t = fmin(t1, t2);
            if (t >= 1e300) { rc = 1; goto done; }   // infeasible
            if (!dependent) {
                prob.step_dir(p, sgp, w.wid, w.wsg, w.r, w.q, z);
                for (int i = lane; i < n; i += 32) x[i] += t * z[i];
                sviol += t * zn;
            }
            for (int k = lane; k < w.q; k += 32) w.mu[k] -= t * w.r[k];
            up += t;
            __syncwarp();
            if (!dependent && t2 <= t1) {
                if (w.q >= w.qmax) { rc = 3; goto done; }
                // y was overwritten?  das_backward only reads y -> still L^-1 c.
                das_append(w, p, sgp, zn, up);
                if (lane == 0) w.state[p] = (signed char)(sgp > 0 ? -1 : +1);
                __syncwarp();
                added = true;
            } else {
                if (lane == 0) w.state[w.wid[l]] = 0;
                __syncwarp();
                das_drop(w, l);
            }
        }
    }
done:
    *iters_out = iters;
    return rc;
}

}  // namespace ismpc
