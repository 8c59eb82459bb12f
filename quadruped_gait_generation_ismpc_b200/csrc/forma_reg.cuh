// forma_reg.cuh -- register-resident build of the structured primal-dual active-set iteration (forma.cuh: forma_pdas).
//
// Same algorithm, same decisions, same per-lane summation order as forma_pdas (trotting/quad_as_bip_bang.m:256 `quadprog`
// replaced by an exact working-set iteration, see the comment block in forma.cuh) -- what changes is where the data
// lives.  forma_pdas walks shared-memory vectors in serial loops with carried state: every row costs a chain of
// dependent shared-memory round trips (state byte -> bound -> prefix sum -> reciprocal table), ~2,500 warp instructions
// and ~14,000 cycles per iteration (ncu / phase clocks, profiles/r1z_forma_tick_*).  Here a lane keeps its E consecutive
// ZMP rows (bounds, stability coefficient and its prefix sum, mapping weight and column, working-set state) in registers
// for the whole solve; the only cross-lane data of an iteration are
//   * the tail (last active row) of the nearest lane below and the head (first active row) of the nearest lane above,
//     fetched with one ballot and a handful of shuffles,
//   * the 14 sums of the saddle system (one halving-payload butterfly),
//   * one shuffle each for the neighbour's state / segment constant at the chunk borders,
// and shared memory is touched only for the reciprocal table, the reduction broadcast and the (rare) general saddle
// solve.  The primal vector is formed once, after the working set has settled.
//
// Covers C <= 32 E rows (E = 4: the scripts' C = 100) and F <= 3 predicted footsteps; other shapes keep forma_pdas.
#pragma once
// (included by forma.cuh, after forma_pdas and its helpers)

namespace ismpc {

template <int E>
__device__ __forceinline__ double sel_row(const double (&v)[E], int le)
{
    double r = v[0];
#pragma unroll
    for (int e = 1; e < E; ++e) r = (le == e) ? v[e] : r;
    return r;
}
template <int E>
__device__ __forceinline__ int sel_row_i(const int (&v)[E], int le)
{
    int r = v[0];
#pragma unroll
    for (int e = 1; e < E; ++e) r = (le == e) ? v[e] : r;
    return r;
}

// On entry sm.lo / sm.hi hold the bounds in SHIFTED coordinates, planf the footstep targets minus shift, sm.das.state
// the starting working set -- the contract of forma_pdas.  On return 0: sm.x = [zd; xf] (xf absolute), sm.rv the row
// values (shifted), sm.das.state the optimal working set.  Returns 0 converged, 1 iteration cap, 2 singular system.
template <int FT, int E>
__device__ __forceinline__ int forma_pdas_reg(const FormAShared& sm, const FormAProb& pb, double beq, double shift,
                                     const double* planf, const double* rg, int maxit, int* iters_out)
{
    static_assert(FT <= 3, "register build: F <= 3 (the 4 x 4 saddle system is solved in registers)");
    constexpr int NS = 2 + 2 * FT + FT * (FT + 1) / 2;
    const int lane = lane_id();
    const int C = pb.C, F = pb.F;
    const double dt = pb.dt, inv_dt = 1.0 / dt, Qz = 1.0 / pb.qz_inv, Qf = 1.0 / pb.qf_inv;
    const double qz_dt = Qz * inv_dt, qz_dt2 = qz_dt * inv_dt, dt_qz = dt * pb.qz_inv;
    signed char* st_s = sm.das.state;        // [C+F] working set (-1 lower, +1 upper)
    double* K = sm.das.Js;                   // scratch: reduction broadcast / augmented system of the general solve
    const int r0 = lane * E;                 // this lane's rows r0 .. r0+E-1 (those below C)
    // ---- row data into registers ----
    // Row constants (lower bound, prefix sum of the stability coefficients, mapping weight / column) are re-read from
    // shared memory at the top of every iteration -- the addresses depend on the lane only, so the loads are issued back
    // to back and their latency hides behind the mask / ballot work; keeping them in registers for the whole solve costs
    // 28 registers and pushed the kernel over the 128 that 16 resident warps per SM allow.  The upper bound of a ZMP row
    // is its lower bound plus the box width, the same for every row of an axis (bang.m:147-150); the stability
    // coefficient a_i is needed by the peeling step and the final primal only.  What stays in registers for the whole
    // solve is the STATE: working set, segment constants, m.xf per row.
    int stv[E];
    const double wbox = pb.hi_[0] - pb.lo_[0];
#pragma unroll
    for (int e = 0; e < E; ++e) { const int i = r0 + e; stv[e] = i < C ? (int)st_s[i] : 0; }
    // coefficient of footstep f in ZMP row i (column f+1 of `mapping`): (p == f+1 ? w : 0) + (p == f ? 1-w : 0).  Tabulated once
    // per solve in the three vectors that are free until the solve ends (x, rv, scr) -- recomputing them from (p, w) with
    // select chains wherever they are used was a quarter of the iteration's instructions.
    double* mct[3] = {sm.x, sm.rv, pb.scr};
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = r0 + e;
        if (i < C) {
            const int p = (int)pb.mp[i]; const double w = pb.mw[i];
#pragma unroll
            for (int f = 0; f < FT; ++f) mct[f][i] = (p == f + 1 ? w : 0.0) + (p == f ? 1.0 - w : 0.0);
        }
    }
    __syncwarp();
#define FORMA_REG_LOAD_ROWS()                                                                          \
    double lo[E], PAv[E], mcv[E][FT];                                                                  \
    _Pragma("unroll") for (int e = 0; e < E; ++e) {                                                    \
        const int ic = r0 + e < C ? r0 + e : C - 1;      /* rows past C mirror the last row; never active, never tested */ \
        lo[e] = pb.lo_[ic]; PAv[e] = pb.PA[ic];                                                        \
        _Pragma("unroll") for (int f = 0; f < FT; ++f) mcv[e][f] = mct[f][ic];                         \
    }                                                                                                  \
    auto hi_of = [&](int e) -> double { return lo[e] + wbox; };                                        \
    auto mc = [&](int e, int f) -> double { return mcv[e][f]; };                                       \
    auto beta = [&](int e) -> double { return stv[e] < 0 ? lo[e] : lo[e] + wbox; };
    int kst = (lane < F) ? (int)st_s[C + lane] : 0;      // kinematic row `lane`
    double qpl[FT];                                      // Qf * footstep target
#pragma unroll
    for (int f = 0; f < FT; ++f) qpl[f] = f < F ? Qf * planf[f] : 0.0;
    double nu = 0.0, xf[FT], cseg[E], mxv[E];
    int kp0 = -1;                            // nearest active row below this lane's chunk, with its bound, prefix sum and m.xf
    double bp0 = 0.0, PAp0 = 0.0, mxp0 = 0.0;
#pragma unroll
    for (int f = 0; f < FT; ++f) xf[f] = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) mxv[e] = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) cseg[e] = 0.0;
    int it = 0, rc = 1;
    DasTimer tmr;
    for (; it < maxit; ++it) {
        tmr.start();
        FORMA_REG_LOAD_ROWS()
        // ---- nearest active rows below / above this lane's chunk ----
        int la = -1, fa = E;
        double t_beta = 0.0, t_PA = 0.0, t_mc[FT];
#pragma unroll
        for (int f = 0; f < FT; ++f) t_mc[f] = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (stv[e]) {
                la = e; if (fa == E) fa = e;
                t_beta = beta(e); t_PA = PAv[e];
#pragma unroll
                for (int f = 0; f < FT; ++f) t_mc[f] = mc(e, f);
            }
        const int nk = __popc(__ballot_sync(ISMPC_FULL_MASK, kst != 0));
        const unsigned has = __ballot_sync(ISMPC_FULL_MASK, la >= 0);
        const unsigned below = has & ((1u << lane) - 1u);
        const unsigned above = lane == 31 ? 0u : has & ~((2u << lane) - 1u);
        const int src_p = below ? 31 - __clz(below) : 0, src_n = above ? __ffs(above) - 1 : 0;
        const int kp_t = __shfl_sync(ISMPC_FULL_MASK, r0 + la, src_p);
        const double bp_t = __shfl_sync(ISMPC_FULL_MASK, t_beta, src_p), PAp_t = __shfl_sync(ISMPC_FULL_MASK, t_PA, src_p);
        double mcp_t[FT];
#pragma unroll
        for (int f = 0; f < FT; ++f) mcp_t[f] = __shfl_sync(ISMPC_FULL_MASK, t_mc[f], src_p);
        kp0 = below ? kp_t : -1;
        bp0 = below ? bp_t : 0.0; PAp0 = below ? PAp_t : 0.0;
        double mcp0[FT];
#pragma unroll
        for (int f = 0; f < FT; ++f) mcp0[f] = below ? mcp_t[f] : 0.0;
        // ---- pass A: sums over the active rows ----
        double acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.0;
        {
            int kp = kp0;
            double bp = bp0, PAp = PAp0, mpr[FT];
#pragma unroll
            for (int f = 0; f < FT; ++f) mpr[f] = mcp0[f];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (!stv[e]) continue;
                const int i = r0 + e;
                const double bk = beta(e), PAk = PAv[e];
                const double w = rg[i - kp], d = PAk - PAp, b = bk - bp;
                const double wd = w * d;
                double ev[FT];
#pragma unroll
                for (int f = 0; f < FT; ++f) { const double mk = mc(e, f); ev[f] = mk - mpr[f]; mpr[f] = mk; }
                acc[0] += wd * d; acc[1] += wd * b;
                int idx = 2 + 2 * FT;
#pragma unroll
                for (int f = 0; f < FT; ++f) {
                    const double we = w * ev[f];
                    acc[2 + f] += d * we; acc[2 + FT + f] += we * b;
#pragma unroll
                    for (int g = 0; g <= f; ++g) { acc[idx] += we * ev[g]; ++idx; }
                }
                kp = i; bp = bk; PAp = PAk;
            }
        }
        tmr.lap(13);
        double kap = 0.0;
        if (nk == 0) {
            // ---- common case: 4 x 4 saddle system solved in registers by every lane (no kinematic row active) ----
            warp_sum_multi16<NS>(acc, K);
            const double k00 = -(pb.saa - acc[0]) * pb.qz_inv, rr0 = -beq + acc[1] * inv_dt;
            double v[3], rf[3], M[3][3];
            int idx = 2 + 2 * FT;
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                v[f] = f < FT ? -acc[2 + (f < FT ? f : 0)] * inv_dt : 0.0;
                rf[f] = f < FT ? qpl[f < FT ? f : 0] - qz_dt2 * acc[2 + FT + (f < FT ? f : 0)] : 0.0;
#pragma unroll
                for (int g = 0; g <= f; ++g) {
                    double m = f == g ? Qf : 0.0;
                    if (f < FT) { m += qz_dt2 * acc[idx < NS ? idx : 0]; ++idx; }
                    M[f][g] = m; M[g][f] = m;
                }
            }
            const double d0 = M[0][0], i0 = fast_rcp(d0);
            const double l10 = M[1][0] * i0, l20 = M[2][0] * i0;
            const double d1 = M[1][1] - l10 * l10 * d0, i1 = fast_rcp(d1);
            const double l21 = (M[2][1] - l20 * l10 * d0) * i1;
            const double d2 = M[2][2] - l20 * l20 * d0 - l21 * l21 * d1, i2 = fast_rcp(d2);
            auto msolve = [&](const double (&b)[3], double (&z)[3]) {
                const double y0 = b[0], y1 = b[1] - l10 * y0, y2 = b[2] - l20 * y0 - l21 * y1;
                z[2] = y2 * i2; z[1] = y1 * i1 - l21 * z[2]; z[0] = y0 * i0 - l10 * z[1] - l20 * z[2];
            };
            double zr[3], zv[3];
            msolve(rf, zr); msolve(v, zv);
            const double S = k00 - (v[0] * zv[0] + v[1] * zv[1] + v[2] * zv[2]);
            if (!(d0 > 0.0) || !(d1 > 0.0) || !(d2 > 0.0) || !(fabs(S) > 1e-200)) { rc = 2; break; }
            nu = (rr0 - (v[0] * zr[0] + v[1] * zr[1] + v[2] * zr[2])) * fast_rcp(S);
#pragma unroll
            for (int f = 0; f < FT; ++f) xf[f] = zr[f] - zv[f] * nu;
        } else {
            // ---- general case: K u = rhs, u = (nu, xf', kappa_active), Gauss-Jordan in shared memory ----
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int s = 0; s < NS; ++s) acc[s] += __shfl_xor_sync(ISMPC_FULL_MASK, acc[s], o);
            }
            const int dim = 1 + F + nk, LD = dim + 1;
            __syncwarp();
            if (lane == 0) {
                for (int s = 0; s < dim * LD; ++s) K[s] = 0.0;
                K[0] = -(pb.saa - acc[0]) * pb.qz_inv;
                K[dim] = -beq + acc[1] * inv_dt;
                int idx = 2 + 2 * FT;
#pragma unroll
                for (int f = 0; f < FT; ++f) {
                    if (f < F) {
                        K[1 + f] = -acc[2 + f] * inv_dt; K[(1 + f) * LD] = -acc[2 + f] * inv_dt;
                        K[(1 + f) * LD + dim] = qpl[f] - qz_dt2 * acc[2 + FT + f];
                    }
#pragma unroll
                    for (int g = 0; g <= f; ++g) {
                        if (f < F) {
                            const double m = qz_dt2 * acc[idx] + (f == g ? Qf : 0.0);
                            K[(1 + f) * LD + 1 + g] = m; K[(1 + g) * LD + 1 + f] = m;
                        }
                        ++idx;
                    }
                }
                int col = 1 + F;
                for (int f = 0; f < F; ++f) {
                    if (!st_s[C + f]) continue;
                    K[(1 + f) * LD + col] -= 1.0; K[col * LD + 1 + f] -= 1.0;
                    if (f > 0) { K[f * LD + col] += 1.0; K[col * LD + f] += 1.0; }
                    K[col * LD + dim] = -(st_s[C + f] < 0 ? pb.lo_[C + f] : pb.hi_[C + f]);
                    ++col;
                }
            }
            __syncwarp();
            if (!warp_gauss_jordan(K, dim)) { rc = 2; break; }
            nu = K[dim] / K[0];
#pragma unroll
            for (int f = 0; f < FT; ++f) xf[f] = f < F ? K[(1 + f) * LD + dim] / K[(1 + f) * LD + 1 + f] : 0.0;
            int col = 1 + F;
            for (int f = 0; f < F; ++f) if (st_s[C + f]) { if (f == lane) kap = K[col * LD + dim] / K[col * LD + col]; ++col; }
            __syncwarp();
        }
        tmr.lap(14);
        // ---- pass B: segment constants, multipliers, row values ----
#pragma unroll
        for (int e = 0; e < E; ++e) {
            double s = 0.0;
#pragma unroll
            for (int f = 0; f < FT; ++f) s += mc(e, f) * xf[f];
            mxv[e] = s;
        }
        mxp0 = 0.0;
#pragma unroll
        for (int f = 0; f < FT; ++f) mxp0 += mcp0[f] * xf[f];
        // (1) constant of the segment that ENDS at each active row
        {
            int kp = kp0;
            double bp = bp0, PAp = PAp0, mxp = mxp0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                cseg[e] = 0.0;
                if (!stv[e]) continue;
                const int i = r0 + e;
                const double bk = beta(e);
                cseg[e] = (qz_dt * ((bk - bp) + (mxv[e] - mxp)) - nu * (PAv[e] - PAp)) * rg[i - kp];
                kp = i; bp = bk; PAp = PAv[e]; mxp = mxv[e];
            }
        }
        // (2) every row takes the constant of the next active row at or after it (0 behind the last one)
        const double head_cin = fa < E ? sel_row<E>(cseg, fa) : 0.0;
        const double cn_t = __shfl_sync(ISMPC_FULL_MASK, head_cin, src_n);
        {
            double run = above ? cn_t : 0.0;
#pragma unroll
            for (int e = E - 1; e >= 0; --e) { if (stv[e]) run = cseg[e]; cseg[e] = run; }
        }
        const double c_dn = __shfl_down_sync(ISMPC_FULL_MASK, cseg[0], 1);
        const double c_after = lane == 31 ? 0.0 : c_dn;                 // constant of the row after this lane's last one
        // (3) row values of the inactive rows, re-guess of the working set
        //     Rows whose multiplier has the wrong sign leave.  From iteration PDAS_DAMP_AFTER on, only those at an end of
        //     a run of equally-signed active rows leave (if there is one): an over-long run can flip the sign of nu,
        //     which would release the whole run at once and make the iteration cycle.
        const int st_up = __shfl_up_sync(ISMPC_FULL_MASK, stv[E - 1], 1), st_dn = __shfl_down_sync(ISMPC_FULL_MASK, stv[0], 1);
        const int st_prev = lane == 0 ? 0 : st_up, st_next = lane == 31 ? 0 : st_dn;
        int s1v[E];
        int changed = 0, wrong_end = 0, wrong_in = 0;
        {
            int kp = kp0;
            double bp = bp0, PAp = PAp0, mxp = mxp0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = r0 + e, s0 = stv[e];
                int s1 = 0;
                if (s0 == 0) {
                    const double r = bp + dt_qz * (nu * (PAv[e] - PAp) + cseg[e] * (double)(i - kp)) - (mxv[e] - mxp);
                    const double hie = hi_of(e);
                    if (i < C) {                                   // (rows past C do not exist)
                        if (lo[e] - r > 1e-10 * (1.0 + fabs(lo[e]))) s1 = -1;
                        else if (r - hie > 1e-10 * (1.0 + fabs(hie))) s1 = +1;
                    }
                } else {
                    const double y = cseg[e] - (e + 1 < E ? cseg[e + 1 < E ? e + 1 : e] : c_after);     // dt * multiplier
                    if (s0 < 0 ? y > 0.0 : y < 0.0) s1 = s0;
                    else {
                        const int sl = e > 0 ? stv[e > 0 ? e - 1 : 0] : st_prev, sr = e + 1 < E ? stv[e + 1 < E ? e + 1 : e] : st_next;
                        s1 = (sl != s0 || sr != s0) ? 0 : 2;          // 2: wrong sign, interior of a run
                        wrong_end |= s1 == 0;
                        wrong_in |= s1 == 2;
                    }
                    bp = beta(e); PAp = PAv[e]; mxp = mxv[e]; kp = i;
                }
                s1v[e] = s1;
            }
        }
        if (lane < F) {
            const int f = lane, s0 = kst;
            double xm1 = 0.0, xme = 0.0;
#pragma unroll
            for (int g = 0; g < FT; ++g) { if (g == f - 1) xm1 = xf[g]; if (g == f) xme = xf[g]; }
            const double r = xme - xm1;
            const double klo = pb.lo_[C + f], khi = pb.hi_[C + f];
            int s1 = 0;
            if (s0 == 0) {
                if (klo - r > 1e-10 * (1.0 + fabs(klo))) s1 = -1;
                else if (r - khi > 1e-10 * (1.0 + fabs(khi))) s1 = +1;
            } else if (s0 < 0 ? kap > 0.0 : kap < 0.0) s1 = s0;
            changed |= s1 != s0;
            sm.rv[C + f] = r;
            sm.x[C + f] = xme + shift;
            st_s[C + f] = (signed char)s1;
            kst = s1;
        }
        tmr.lap(15);
        const unsigned end_mask = __ballot_sync(ISMPC_FULL_MASK, wrong_end);
        // ---- peeling step (see forma_pdas): a run whose end row has a wrong-sign multiplier is cut back in one step to
        // the first row whose multiplier keeps its sign, with nu and the footsteps frozen ----
#ifdef ISMPC_FORMA_PEEL_CALL
        if (end_mask != 0u && !__any_sync(ISMPC_FULL_MASK, wrong_in)) {
            // out of line, on shared-memory copies of the state (forma_peel_smem): stage, call, take the cuts back
            int* nxt = sm.das.wid;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = r0 + e;
                if (i < C) { st_s[i] = (signed char)stv[e]; nxt[i] = s1v[e]; pb.scr[i] = cseg[e]; }
            }
            __syncwarp();
            forma_peel_smem(&pb, st_s, nxt, pb.scr, rg, nu, xf[0], FT > 1 ? xf[FT > 1 ? 1 : 0] : 0.0, FT > 2 ? xf[FT > 2 ? 2 : 0] : 0.0,
                            end_mask, qz_dt);
            __syncwarp();
#pragma unroll
            for (int e = 0; e < E; ++e) { const int i = r0 + e; if (i < C) s1v[e] = nxt[i]; }
        }
#else
        if (end_mask != 0u && !__any_sync(ISMPC_FULL_MASK, wrong_in)) {
            unsigned todo = end_mask;
            while (todo) {
                const int L = __ffs(todo) - 1; todo &= todo - 1;
#pragma unroll 1
                for (int le = 0; le < E; ++le) {
                    const int i = L * E + le;
                    // state and staged state of row i, as they are NOW (an earlier cut may have marked it)
                    const int sg = __shfl_sync(ISMPC_FULL_MASK, sel_row_i<E>(stv, le), L);
                    const int s1i = __shfl_sync(ISMPC_FULL_MASK, sel_row_i<E>(s1v, le), L);
                    if (sg == 0 || s1i != 0) continue;                 // not a wrong end row (3 = cut by an earlier end)
                    const int sgn_r = __shfl_sync(ISMPC_FULL_MASK, le + 1 < E ? sel_row_i<E>(stv, le + 1 < E ? le + 1 : le) : st_next, L);
                    const int sgn_l = __shfl_sync(ISMPC_FULL_MASK, le > 0 ? sel_row_i<E>(stv, le > 0 ? le - 1 : 0) : st_prev, L);
                    const bool right = i + 1 >= C || sgn_r != sg, left = i == 0 || sgn_l != sg;
                    // targets of this lane's rows as members of a run of sign sg
                    double tg[E];
#pragma unroll
                    for (int e = 0; e < E; ++e) tg[e] = (sg < 0 ? lo[e] : lo[e] + wbox) + mxv[e];
                    if (right) {
                        int lb = -1, kn = C;                           // last row before i outside the run, next active row after i
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            const int k = r0 + e;
                            if (k < i && k < C && stv[e] != sg) lb = k;
                            if (k > i && stv[e] != 0 && kn == C) kn = k;
                        }
                        lb = __reduce_max_sync(ISMPC_FULL_MASK, lb); kn = __reduce_min_sync(ISMPC_FULL_MASK, kn);
                        const int s = lb + 1;
                        const int knl = kn < C ? kn / E : 0, kne = kn < C ? kn - knl * E : 0;
                        double own[E];
#pragma unroll
                        for (int e = 0; e < E; ++e) own[e] = beta(e) + mxv[e];
                        const double tkn_b = __shfl_sync(ISMPC_FULL_MASK, sel_row<E>(own, kne), knl);
                        const double PAkn_b = __shfl_sync(ISMPC_FULL_MASK, sel_row<E>(PAv, kne), knl);
                        const double tkn = kn < C ? tkn_b : 0.0, PAkn = kn < C ? PAkn_b : 0.0;
                        const double tg_up = __shfl_up_sync(ISMPC_FULL_MASK, tg[E - 1], 1);      // target of row r0-1 (in the run when used)
                        int best = -1;
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            const int k = r0 + e;
                            if (k < s || k > i) continue;
                            const double te = tg[e];
                            const double c_new = kn < C ? (qz_dt * (tkn - te) - nu * (PAkn - PAv[e])) * rg[kn - k] : 0.0;
                            const double c_prev = k > s ? qz_dt * (te - (e > 0 ? tg[e > 0 ? e - 1 : 0] : tg_up)) - nu * pb.a[k] : cseg[e];
                            const double y = c_prev - c_new;
                            if (sg < 0 ? y > 0.0 : y < 0.0) best = k;
                        }
                        best = __reduce_max_sync(ISMPC_FULL_MASK, best);
                        const int from = best >= s ? best + 1 : s;
#pragma unroll
                        for (int e = 0; e < E; ++e) { const int k = r0 + e; if (k >= from && k <= i) s1v[e] = 3; }
                    }
                    if (left) {
                        int ub = C, kp = -1;                           // first row after i outside the run, last active row before i
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            const int k = r0 + e;
                            if (k > i && k < C && stv[e] != sg && ub == C) ub = k;
                            if (k < i && stv[e] != 0) kp = k;
                        }
                        ub = __reduce_min_sync(ISMPC_FULL_MASK, ub); kp = __reduce_max_sync(ISMPC_FULL_MASK, kp);
                        const int ee = ub - 1;
                        const int kpl = kp >= 0 ? kp / E : 0, kpe = kp >= 0 ? kp - kpl * E : 0;
                        double own[E];
#pragma unroll
                        for (int e = 0; e < E; ++e) own[e] = beta(e) + mxv[e];
                        const double tkp_b = __shfl_sync(ISMPC_FULL_MASK, sel_row<E>(own, kpe), kpl);
                        const double PAkp_b = __shfl_sync(ISMPC_FULL_MASK, sel_row<E>(PAv, kpe), kpl);
                        const double tkp = kp >= 0 ? tkp_b : 0.0, PAkp = kp >= 0 ? PAkp_b : 0.0;
                        const int al = ee + 1 < C ? (ee + 1) / E : 0, ae = ee + 1 < C ? (ee + 1) - al * E : 0;
                        const double ca_b = __shfl_sync(ISMPC_FULL_MASK, sel_row<E>(cseg, ae), al);
                        const double c_aft = ee + 1 < C ? ca_b : 0.0;
                        const double tg_dn = __shfl_down_sync(ISMPC_FULL_MASK, tg[0], 1);
                        int best = C;
#pragma unroll
                        for (int e = E - 1; e >= 0; --e) {
                            const int k = r0 + e;
                            if (k < i || k > ee) continue;
                            const double ts = tg[e];
                            const double c_new = (qz_dt * (ts - tkp) - nu * (PAv[e] - PAkp)) * rg[k - kp];
                            const double t_nx = e + 1 < E ? tg[e + 1 < E ? e + 1 : e] : tg_dn, a_nx = k + 1 < C ? pb.a[k + 1 < C ? k + 1 : k] : 0.0;
                            const double c_next = k < ee ? qz_dt * (t_nx - ts) - nu * a_nx : c_aft;
                            const double y = c_new - c_next;
                            if (sg < 0 ? y > 0.0 : y < 0.0) best = k;
                        }
                        best = __reduce_min_sync(ISMPC_FULL_MASK, best);
                        const int to = best <= ee ? best - 1 : ee;
#pragma unroll
                        for (int e = 0; e < E; ++e) { const int k = r0 + e; if (k >= i && k <= to) s1v[e] = 3; }
                    }
                }
            }
        }
#endif
        {
            const bool damp = it >= PDAS_DAMP_AFTER && end_mask != 0u;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                int s1 = s1v[e];
                if (s1 == 2) s1 = damp ? stv[e] : 0;
                if (s1 == 3) s1 = 0;
                changed |= s1 != stv[e];
                stv[e] = s1;
            }
        }
        changed = __any_sync(ISMPC_FULL_MASK, changed);
        tmr.lap(16);
        if (!changed) { rc = 0; ++it; break; }
    }
    // ---- write back: working set, and (for the settled set) the primal vector and the row values ----
    // (stv is the set the last iteration solved for: nothing changed in it, so nu, xf, cseg and the predecessor data
    // of that iteration belong to it)
    {
        FORMA_REG_LOAD_ROWS()
        (void)hi_of; (void)mc;
        int kp = kp0;
        double bp = bp0, PAp = PAp0, mxp = mxp0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = r0 + e;
            if (i >= C) continue;
            double r;
            if (stv[e]) { r = beta(e); bp = r; PAp = PAv[e]; mxp = mxv[e]; kp = i; }
            else r = bp + dt_qz * (nu * (PAv[e] - PAp) + cseg[e] * (double)(i - kp)) - (mxv[e] - mxp);
            sm.x[i] = (nu * pb.a[i] + cseg[e]) * pb.qz_inv;
            sm.rv[i] = r;
            st_s[i] = (signed char)stv[e];
        }
    }
    __syncwarp();
#undef FORMA_REG_LOAD_ROWS
    *iters_out = it;
    return rc;
}

}  // namespace ismpc
